#!/usr/bin/env bash
# cfg C joins the full-size parity set (every BASELINE.json configuration at real size and depth)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_fullsize_gpu.py -q -m gpu -p no:cacheprovider --tb=short -rf -s -k "forward" 2>&1 | grep -v Warning | tail -14
