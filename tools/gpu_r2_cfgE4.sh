#!/usr/bin/env bash
# BASELINE.json config 5: SDEdit upsampling 1024^2, 20 steps, batch 4 over 4 B200 (one sample per GPU) + extra.sp
mkdir -p gpurun_out
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29631 bench.py --gpus 4 --workload E --steps 5 --warmup 3 > gpurun_out/r2_bench_E_n4.json 2> gpurun_out/r2_bench_E_n4.err
tail -c 600 gpurun_out/r2_bench_E_n4.err; tail -c 1200 gpurun_out/r2_bench_E_n4.json
