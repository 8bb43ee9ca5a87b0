#!/usr/bin/env bash
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:attn_fwd2_tcgen05 -s 3 -c 1 -f -o gpurun_out/prof_attn2 python tools/time_full.py 2 > gpurun_out/ncu_attn2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:ln_modulate -s 8 -c 2 -f -o gpurun_out/prof_ln python tools/time_full.py 2 > gpurun_out/ncu_ln.log 2>&1
ncu --set full --clock-control none --import-source on -k "regex:gemm_bf16_tcgen05_kernel<256, 2, [34]" -s 40 -c 3 -f -o gpurun_out/prof_gemm_qkv python tools/time_full.py 2 > gpurun_out/ncu_gemm_qkv.log 2>&1
ls -la gpurun_out/*.ncu-rep
