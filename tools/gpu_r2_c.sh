#!/usr/bin/env bash
# round-2 GPU session C: whole -m gpu suite on the current defaults, attention burst + sustained vs libraries, ncu evidence
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider 2>&1 | tail -40 > gpurun_out/r2c_pytest.log
tail -8 gpurun_out/r2c_pytest.log
timeout 400 python tools/bench_attn_libs.py > gpurun_out/r2c_attn_libs.log 2>&1
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
bash tools/gpu_r2_profile.sh
