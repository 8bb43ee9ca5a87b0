#!/usr/bin/env bash
# Runs GPU test groups one by one, each under its own timeout, so one hung kernel cannot hide the others.
# Usage (on the GPU box, from the repo root): tools/gpu_diag.sh [pytest -k expressions...]
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.used --format=csv > gpurun_out/diag_gpu.txt 2>&1
GROUPS_DEFAULT=("probe" "gemm_bias and (0-1 or 128-1 or 192-1 or 256-1)" "gemm_bias and (0-2 or 128-2 or 192-2 or 256-2)" "gemm_block64 or gemm_gelu or gemm_gate" "gemm_qkv or gemm_linear1" "attention" "ln_modulate or timestep or rope_table")
if [[ $# -gt 0 ]]; then GROUPS_SEL=("$@"); else GROUPS_SEL=("${GROUPS_DEFAULT[@]}"); fi
i=0
for g in "${GROUPS_SEL[@]}"; do
  i=$((i+1))
  echo "=== group $i: $g"
  timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q --no-header -p no:cacheprovider -k "$g" > "gpurun_out/diag_$i.log" 2>&1
  echo "exit=$?" >> "gpurun_out/diag_$i.log"
  tail -n 25 "gpurun_out/diag_$i.log" | cut -c1-400
done
