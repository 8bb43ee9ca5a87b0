#!/usr/bin/env bash
# BASELINE.json config 3: 512 grid 2x3, 30 steps, batch 8 sharded over 8 B200 (one sample per GPU) + extra.sp (one image over 8)
mkdir -p gpurun_out
timeout 700 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29621 bench.py --gpus 8 --workload C --steps 5 --warmup 3 > gpurun_out/r2_bench_C_n8.json 2> gpurun_out/r2_bench_C_n8.err
tail -c 600 gpurun_out/r2_bench_C_n8.err; tail -c 1200 gpurun_out/r2_bench_C_n8.json
