#!/usr/bin/env bash
# round-2 closing validation on one GPU with the final library: smoke(), the whole -m gpu suite, the four bench lines + the reference
# arm, and ncu captures of the two kernels fp8 level 2 added
mkdir -p gpurun_out; rm -f gpurun_out/fp8_parity.json gpurun_out/fullsize_parity.json
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider --tb=short -rf 2>&1 | grep -v Warning | tail -30 > gpurun_out/r2n_pytest.log
tail -5 gpurun_out/r2n_pytest.log
show() { python - "$1" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], {k: d[k] for k in ("value", "ms_per_step", "dtype")}, "e2e", d["e2e"]["value"], d["clocks"]["sm_mhz"])
    for k in ("roofline", "roofline_attention", "roofline_ln_modulate", "roofline_vae"):
        print(" ", k, round(d[k]["achieved"], 1), round(d[k]["frac"], 3))
    print(" ", d["kernel_time_share"]); print(" ", d["extra"]); print(" ", d.get("cpu_baseline"))
except Exception as e:
    print("bench parse failed", sys.argv[1], e)
PY
}
timeout 900 python bench.py > gpurun_out/r2n_bench.json 2> gpurun_out/r2n_bench.err; tail -c 300 gpurun_out/r2n_bench.err; show gpurun_out/r2n_bench.json
timeout 900 python bench.py --precision fp8 --no-cpu-baseline > gpurun_out/r2n_bench_fp8.json 2> gpurun_out/r2n_bench_fp8.err; show gpurun_out/r2n_bench_fp8.json
timeout 900 python bench.py --precision fp8_all --no-cpu-baseline > gpurun_out/r2n_bench_fp8_all.json 2> gpurun_out/r2n_bench_fp8_all.err; show gpurun_out/r2n_bench_fp8_all.json
timeout 600 python bench.py --impl reference > gpurun_out/r2n_bench_reference.json 2> gpurun_out/r2n_bench_reference.err; tail -c 600 gpurun_out/r2n_bench_reference.json
cap() {  # name, kernel regex
  ncu --set full --clock-control none --import-source on -k "regex:$2" -s 1 -c 1 -f -o "gpurun_out/r2_prof_$1" python tools/ncu_targets.py "$1" > "gpurun_out/r2_ncu_$1.log" 2>&1
  tail -2 "gpurun_out/r2_ncu_$1.log"
  if [ -f "gpurun_out/r2_prof_$1.ncu-rep" ]; then
    ncu -i "gpurun_out/r2_prof_$1.ncu-rep" --page raw --csv > "gpurun_out/r2_prof_$1.raw.csv" 2>/dev/null
    rm -f "gpurun_out/r2_prof_$1.ncu-rep"
  fi
}
cap quantize_cat quantize_rows_e4m3
cap gemm_fp8_linear2 gemm_bf16_tcgen05
cap gemm_fp8_linear1 gemm_bf16_tcgen05
ls -la gpurun_out/r2_prof_*.raw.csv | tail -4
