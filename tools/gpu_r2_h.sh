#!/usr/bin/env bash
# round-2 GPU session H: source-level ncu captures of the per-pair and the persistent attention kernel (reports come back: 2 x 16 MB)
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep
for t in attn_pair attn_persistent; do
  k=attn_fwd3; [ $t = attn_persistent ] && k=attn_fwd4
  ncu --set full --clock-control none --import-source on -k "regex:$k" -s 1 -c 1 -f -o "gpurun_out/r2h_$t" python tools/ncu_targets.py $t > "gpurun_out/r2h_ncu_$t.log" 2>&1
  tail -1 "gpurun_out/r2h_ncu_$t.log"
done
ls -la gpurun_out/*.ncu-rep; du -sh gpurun_out
