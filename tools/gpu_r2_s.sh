#!/usr/bin/env bash
# native text encoders: kernel + model parity against the HF modules, then the once-per-image cost at the real geometry
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_text_encoders_gpu.py -q -m gpu -p no:cacheprovider --tb=short -rf -s 2>&1 | grep -v Warning | tail -40 > gpurun_out/r2s_pytest.log
tail -30 gpurun_out/r2s_pytest.log
timeout 600 python tools/bench_text_encoders.py 2>&1 | tail -4
