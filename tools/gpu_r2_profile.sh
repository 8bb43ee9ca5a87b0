#!/usr/bin/env bash
# round-2 ncu evidence (one GPU; never a number taken from these runs is a bench value):
#  (1) launch list of prepare + 2 evaluations of cfg B, (2) one --set full capture per hot kernel at its cfg-B shape
mkdir -p gpurun_out
# only this library's kernels (the model's random init alone launches > 1200 torch fill / RNG kernels before the first forward)
OURS='regex:tcgen05|ln_modulate|euler_update|rope_table|timestep_embedding|silu_kernel|add3_kernel|copy_cols|gn_|softmax_rows|tokens_to_nhwc|nhwc_to_image|upsample2x|moments_to_tokens|sp_barrier'
ncu --metrics gpu__time_duration.sum --clock-control none -k "$OURS" -c 1400 --csv --log-file gpurun_out/r2_launches.csv python tools/time_full.py 3 > gpurun_out/r2_launches.log 2>&1
cap() {  # name, kernel regex
  ncu --set full --clock-control none --import-source on -k "regex:$2" -s 1 -c 1 -f -o "gpurun_out/r2_prof_$1" python tools/ncu_targets.py "$1" > "gpurun_out/r2_ncu_$1.log" 2>&1
  tail -2 "gpurun_out/r2_ncu_$1.log"
  # the .ncu-rep files are ~16 MB each and gpurun_out/ is capped at 64 MiB: keep the raw-page CSV (all metrics of the launch)
  if [ -f "gpurun_out/r2_prof_$1.ncu-rep" ]; then
    ncu -i "gpurun_out/r2_prof_$1.ncu-rep" --page raw --csv > "gpurun_out/r2_prof_$1.raw.csv" 2>/dev/null
    rm -f "gpurun_out/r2_prof_$1.ncu-rep"
  fi
}
cap ln ln_modulate2
cap attn_pair attn_fwd3
cap attn_pair_exact attn_fwd3
cap attn_persistent attn_fwd4
cap gemm_linear1 gemm_bf16_tcgen05
cap gemm_linear2 gemm_bf16_tcgen05
cap gemm_fp8_linear1 gemm_bf16_tcgen05
cap conv512 gemm_bf16_tcgen05
cap conv256 gemm_bf16_tcgen05
cap conv128 gemm_bf16_tcgen05
ls -la gpurun_out/r2_prof_*.raw.csv
