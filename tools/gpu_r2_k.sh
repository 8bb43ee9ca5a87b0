#!/usr/bin/env bash
# fp8 level 2 ("fp8_all": proj / mlp.2 / linear2 on e4m3 too): kernel tests, the parity study, the three bench lines
mkdir -p gpurun_out; rm -f gpurun_out/fp8_parity.json
timeout 1200 python -m pytest tests/test_fp8_gpu.py -q -m gpu -p no:cacheprovider --tb=short -rf -s 2>&1 | grep -v Warning | tail -40 > gpurun_out/r2k_pytest_fp8.log
tail -22 gpurun_out/r2k_pytest_fp8.log
for prec in fp8_all fp8; do
  timeout 900 python bench.py --precision $prec --no-cpu-baseline > gpurun_out/r2k_bench_$prec.json 2> gpurun_out/r2k_bench_$prec.err; tail -c 300 gpurun_out/r2k_bench_$prec.err
  python - $prec <<'PY'
import json, sys
try:
    d = json.loads(open(f"gpurun_out/r2k_bench_{sys.argv[1]}.json").read().strip().splitlines()[-1])
    print(sys.argv[1], {k: d[k] for k in ("value", "ms_per_step")}, "e2e", d["e2e"]["value"], d["clocks"]["sm_mhz"]); print(" ", d["kernel_time_share"])
    for r in d["gemm_shapes"]: print("   ", r["M"], r["N"], r["K"], r["epilogue"], r["block_n"], r["cta_group"], r["launches"], round(r["avg_us"], 1), round(r["tflops"]))
except Exception as e:
    print("bench parse failed", e)
PY
done
timeout 900 python -m pytest tests/test_flux_gpu.py tests/test_kernels_gpu.py -q -m gpu -p no:cacheprovider -x 2>&1 | tail -3
