#!/usr/bin/env bash
# round-2 GPU session B: rescheduled persistent attention, LN with fp32 staging, fp8 path, per-shape GEMM table, stream-K A/B
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -p no:cacheprovider -k "attention or ln" 2>&1 | tail -30 > gpurun_out/r2b_kernel_tests.log
tail -6 gpurun_out/r2b_kernel_tests.log
timeout 900 python -m pytest tests/test_fp8_gpu.py -q -p no:cacheprovider -s 2>&1 | tail -60 > gpurun_out/r2b_fp8_tests.log
tail -25 gpurun_out/r2b_fp8_tests.log
timeout 300 python tools/bench_attn_libs.py > gpurun_out/r2b_attn_libs.log 2>&1
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/attn_vs_libs.json"))
    for r in d["rows"]:
        print(r["L"], {k: round(v["tflops"]) for k, v in r.items() if isinstance(v, dict) and "tflops" in v})
except Exception as e:
    print("attn libs parse failed", e)
PY
show() { python - "$1" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], {k: d[k] for k in ("value", "ms_per_step", "dtype")}, "e2e", d["e2e"]["value"], d["clocks"]["sm_mhz"])
    for k in ("roofline", "roofline_attention", "roofline_ln_modulate", "roofline_vae"):
        print(" ", k, round(d[k]["achieved"], 1), round(d[k]["frac"], 3))
    print(" ", d["kernel_time_share"]); print(" ", d["extra"])
    for r in d["gemm_shapes"]: print("   ", {k: (round(v, 3) if isinstance(v, float) else v) for k, v in r.items()})
except Exception as e:
    print("bench parse failed", sys.argv[1], e)
PY
}
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err; tail -c 800 gpurun_out/r2b_bench.err; show gpurun_out/r2b_bench.json
timeout 900 python bench.py --steps 3 --warmup 3 --precision fp8 --no-cpu-baseline > gpurun_out/r2b_bench_fp8.json 2> gpurun_out/r2b_bench_fp8.err; tail -c 800 gpurun_out/r2b_bench_fp8.err; show gpurun_out/r2b_bench_fp8.json
timeout 300 python tools/gemm_shape_table.py > gpurun_out/r2b_gemm_shapes.log 2>&1; tail -8 gpurun_out/r2b_gemm_shapes.log
VCB_STREAMK=1 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2b_bench_streamk.json 2> gpurun_out/r2b_bench_streamk.err; show gpurun_out/r2b_bench_streamk.json
