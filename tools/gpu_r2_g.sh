#!/usr/bin/env bash
# round-2 GPU session G: where the persistent attention kernel loses its time (timelines), after removing local-memory traffic
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -q -p no:cacheprovider --tb=short -k "attention" 2>&1 | tail -4
for L in 3968 6656; do
  VCB_ATTN4_TIMELINE=1 timeout 120 python tools/attn4_timeline.py $L 2>&1 | tail -12
  VCB_ATTN4_TIMELINE=1 VCB_ATTN4_NOSPLIT=1 timeout 120 python tools/attn4_timeline.py $L 2>&1 | tail -12
done
VCB_ATTN4_TIMELINE=1 timeout 120 python tools/attn4_timeline.py 3968 6 2>&1 | tail -12
