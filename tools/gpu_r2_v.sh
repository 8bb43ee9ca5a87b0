#!/usr/bin/env bash
# sequence-parallel parity at 4 ranks with the final kernels (the world=4 case is skipped on smaller boxes)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_sp_gpu.py -q -p no:cacheprovider -rs 2>&1 | tail -6
