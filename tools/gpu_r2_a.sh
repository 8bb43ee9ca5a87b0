#!/usr/bin/env bash
# round-2 GPU session A: persistent attention + two-pass LayerNorm correctness, library comparison, full-size parity
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -p no:cacheprovider -k "attention or ln_modulate or layernorm or ln" 2>&1 | tail -40 > gpurun_out/r2a_kernel_tests.log
tail -15 gpurun_out/r2a_kernel_tests.log
timeout 300 python tools/bench_attn_libs.py > gpurun_out/r2a_attn_libs.log 2>&1
tail -8 gpurun_out/r2a_attn_libs.log
VCB_ATTN_PERSIST=0 timeout 900 python -m pytest tests/test_fullsize_gpu.py -q -p no:cacheprovider -s 2>&1 | tail -60 > gpurun_out/r2a_fullsize.log
tail -12 gpurun_out/r2a_fullsize.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err
tail -c 1500 gpurun_out/r2a_bench.err; python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r2a_bench.json").read().strip().splitlines()[-1])
    print({k: d[k] for k in ("value", "ms_per_step", "clocks")}, d["e2e"]["value"])
    for k in ("roofline", "roofline_attention", "roofline_ln_modulate", "roofline_vae"):
        print(k, round(d[k]["achieved"], 1), round(d[k]["frac"], 3))
    print(d["kernel_time_share"]); print(d["extra"])
    for r in d["gemm_shapes"]: print(r)
except Exception as e:
    print("bench parse failed", e)
PY
