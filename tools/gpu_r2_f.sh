#!/usr/bin/env bash
# round-2 GPU session F: N-fastest tile raster for the A-heavy GEMMs -- tests, in-loop A/B, ncu DRAM traffic of linear2
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_flux_gpu.py -q -p no:cacheprovider --tb=short -rf -k "gemm or ln or raster or forward or sampler" 2>&1 | tail -25 > gpurun_out/r2f_tests.log
tail -6 gpurun_out/r2f_tests.log
show() { python - "$1" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], {k: d[k] for k in ("value", "ms_per_step")}, "e2e", d["e2e"]["value"], d["clocks"]["sm_mhz"])
    for k in ("roofline", "roofline_attention", "roofline_ln_modulate"):
        print(" ", k, round(d[k]["achieved"], 1), round(d[k]["frac"], 3))
    print(" ", d["kernel_time_share"])
    for r in d["gemm_shapes"]: print("   ", {k: (round(v, 3) if isinstance(v, float) else v) for k, v in r.items()})
except Exception as e:
    print("bench parse failed", sys.argv[1], e)
PY
}
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err; tail -c 400 gpurun_out/r2f_bench.err; show gpurun_out/r2f_bench.json
VCB_GEMM_RASTER=0 timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2f_bench_mfast.json 2> gpurun_out/r2f_bench_mfast.err; show gpurun_out/r2f_bench_mfast.json
ncu --set full --clock-control none -k regex:gemm_bf16_tcgen05 -s 1 -c 1 -f -o gpurun_out/r2_prof_gemm_linear2_nfast python tools/ncu_targets.py gemm_linear2 > gpurun_out/r2_ncu_gemm_linear2_nfast.log 2>&1
ncu -i gpurun_out/r2_prof_gemm_linear2_nfast.ncu-rep --page raw --csv > gpurun_out/r2_prof_gemm_linear2_nfast.raw.csv 2>/dev/null; rm -f gpurun_out/r2_prof_gemm_linear2_nfast.ncu-rep
python - <<'PY'
import csv
rows = list(csv.reader(open("gpurun_out/r2_prof_gemm_linear2_nfast.raw.csv")))
h, u = rows[0], rows[1]
for r in rows[2:]:
    for n in ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct"):
        if n in h: print(n, r[h.index(n)], u[h.index(n)])
PY
