#!/usr/bin/env bash
# round-2 GPU session E: LayerNorm statistics from the GEMM epilogue -- whole suite, bench lines, ncu of the new LN + launch list
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider --tb=short -rf 2>&1 | tail -60 > gpurun_out/r2e_pytest.log
tail -12 gpurun_out/r2e_pytest.log
show() { python - "$1" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], {k: d[k] for k in ("value", "ms_per_step", "dtype")}, "e2e", d["e2e"]["value"], d["clocks"]["sm_mhz"])
    for k in ("roofline", "roofline_attention", "roofline_ln_modulate", "roofline_vae"):
        print(" ", k, round(d[k]["achieved"], 1), round(d[k]["frac"], 3))
    print(" ", d["kernel_time_share"]); print(" ", d["extra"])
except Exception as e:
    print("bench parse failed", sys.argv[1], e)
PY
}
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.err; tail -c 600 gpurun_out/r2e_bench.err; show gpurun_out/r2e_bench.json
timeout 900 python bench.py --steps 3 --warmup 3 --precision fp8 --no-cpu-baseline > gpurun_out/r2e_bench_fp8.json 2> gpurun_out/r2e_bench_fp8.err; show gpurun_out/r2e_bench_fp8.json
timeout 900 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/r2e_bench_reference.json 2> gpurun_out/r2e_bench_reference.err; tail -c 700 gpurun_out/r2e_bench_reference.json
OURS='regex:tcgen05|ln_modulate|euler_update|rope_table|timestep_embedding|silu_kernel|add3_kernel|copy_cols|gn_|softmax_rows|tokens_to_nhwc|nhwc_to_image|upsample2x|moments_to_tokens|sp_barrier'
ncu --metrics gpu__time_duration.sum --clock-control none -k "$OURS" -c 1400 --csv --log-file gpurun_out/r2_launches.csv python tools/time_full.py 3 > gpurun_out/r2_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:ln_modulate_stats -s 1 -c 1 -f -o gpurun_out/r2_prof_ln_stats python tools/ncu_targets.py ln_stats > gpurun_out/r2_ncu_ln_stats.log 2>&1
ncu -i gpurun_out/r2_prof_ln_stats.ncu-rep --page raw --csv > gpurun_out/r2_prof_ln_stats.raw.csv 2>/dev/null; rm -f gpurun_out/r2_prof_ln_stats.ncu-rep
du -sh gpurun_out
