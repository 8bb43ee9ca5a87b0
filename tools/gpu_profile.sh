#!/usr/bin/env bash
# ncu evidence for the round: (1) launch list with device time of every vcb kernel over 2 model evaluations,
# (2) full-set capture of the dominant GEMM and of the attention kernel.  Run on the GPU box from the repo root.
mkdir -p gpurun_out
KREGEX='regex:gemm_bf16_tcgen05|attn_fwd|ln_modulate|euler_update|silu_kernel|add3|rope_table|timestep_embedding|copy_cols|vae_|gn_|conv'
ncu --metrics gpu__time_duration.sum --clock-control none -k "$KREGEX" -c 1200 --csv --log-file gpurun_out/launches.csv \
    python tools/time_full.py 3 > gpurun_out/ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_tcgen05 -s 200 -c 4 -f -o gpurun_out/prof_gemm \
    python tools/time_full.py 2 > gpurun_out/ncu_gemm.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:attn_fwd -s 3 -c 2 -f -o gpurun_out/prof_attn \
    python tools/time_full.py 2 > gpurun_out/ncu_attn.log 2>&1
ls -la gpurun_out | head -30
