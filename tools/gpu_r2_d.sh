#!/usr/bin/env bash
# round-2 GPU session D: whole -m gpu suite with short tracebacks, attention burst + sustained vs libraries, ncu evidence (CSV only)
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep
timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider --tb=short -rf 2>&1 | tail -150 > gpurun_out/r2d_pytest.log
tail -25 gpurun_out/r2d_pytest.log
timeout 400 python tools/bench_attn_libs.py > gpurun_out/r2d_attn_libs.log 2>&1
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
du -sh gpurun_out
