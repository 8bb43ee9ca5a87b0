#!/usr/bin/env bash
# round-2 GPU session I: persistent attention after the wait-loop / spill fixes -- tests, timeline, burst + sustained vs libraries
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -q -p no:cacheprovider --tb=short -k "attention" 2>&1 | tail -4
VCB_ATTN4_TIMELINE=1 timeout 120 python tools/attn4_timeline.py 3968 2>&1 | tail -9
VCB_ATTN4_TIMELINE=1 timeout 120 python tools/attn4_timeline.py 6656 2>&1 | tail -9
timeout 400 python tools/bench_attn_libs.py > gpurun_out/r2i_attn_libs.log 2>&1
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/attn_vs_libs.json"))
    for r in d["rows"]:
        print(r["L"], "burst", {k: round(v["tflops"]) for k, v in r.items() if isinstance(v, dict) and "tflops" in v})
        print(r["L"], "sustained", {k: round(v["sustained_tflops"]) for k, v in r.items() if isinstance(v, dict) and "sustained_tflops" in v})
except Exception as e:
    print("attn libs parse failed", e)
PY
