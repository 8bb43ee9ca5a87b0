#!/usr/bin/env bash
# full-set captures of specific GEMM launches by index inside `tools/time_full.py 2` (prepare = 84 GEMMs, then forward):
# 85 img qkv, 87 img proj, 88 img mlp-up, 89 img mlp-down, 237 linear1, 238 linear2
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_tcgen05 -s 85 -c 5 -f -o gpurun_out/prof_gemm_double python tools/time_full.py 2 > gpurun_out/ncu_gd.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_tcgen05 -s 237 -c 2 -f -o gpurun_out/prof_gemm_single python tools/time_full.py 2 > gpurun_out/ncu_gs.log 2>&1
