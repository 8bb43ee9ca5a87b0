#!/usr/bin/env bash
# A/B of the epilogue restructuring (RoPE prefetch, hoisted loads, fp8 without intermediate roundings): probe old vs new library,
# then the new library in place of the in-tree one (box copy only): kernel / fp8 / model tests and the fp8_all + bf16 bench lines
mkdir -p gpurun_out
python tools/fp8_gemm_probe.py > gpurun_out/r2m_probe_old.log 2>&1; tail -7 gpurun_out/r2m_probe_old.log | head -6
cp tools/_probe/libvcb200_new.so visualcloze_b200/libvcb200.so
python tools/fp8_gemm_probe.py > gpurun_out/r2m_probe_new.log 2>&1; tail -7 gpurun_out/r2m_probe_new.log | head -6
timeout 1200 python -m pytest tests/test_kernels_gpu.py tests/test_fp8_gpu.py tests/test_flux_gpu.py tests/test_fullsize_gpu.py -q -m gpu -p no:cacheprovider --tb=short -rf -x 2>&1 | grep -v Warning | tail -15 > gpurun_out/r2m_pytest.log
tail -8 gpurun_out/r2m_pytest.log
for prec in fp8_all bf16; do
  timeout 900 python bench.py --precision $prec --no-cpu-baseline > gpurun_out/r2m_bench_$prec.json 2> gpurun_out/r2m_bench_$prec.err; tail -c 300 gpurun_out/r2m_bench_$prec.err
  python - $prec <<'PY'
import json, sys
try:
    d = json.loads(open(f"gpurun_out/r2m_bench_{sys.argv[1]}.json").read().strip().splitlines()[-1])
    print(sys.argv[1], {k: d[k] for k in ("value", "ms_per_step")}, "e2e", d["e2e"]["value"], d["clocks"]["sm_mhz"]); print(" ", d["kernel_time_share"])
    for r in d["gemm_shapes"]: print("   ", r["M"], r["N"], r["K"], r["epilogue"], r["launches"], round(r["avg_us"], 1), round(r["tflops"]))
except Exception as e:
    print("bench parse failed", e)
PY
done
