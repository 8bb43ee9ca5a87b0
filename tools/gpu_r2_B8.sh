#!/usr/bin/env bash
# the contract's cfg-B line at N = 8 (replicas) + extra.sp: ONE cfg-B image over 8 GPUs (persistent attention: 48 per-pair CTAs)
mkdir -p gpurun_out
timeout 700 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29641 bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/r2_bench_B_n8.json 2> gpurun_out/r2_bench_B_n8.err
tail -c 500 gpurun_out/r2_bench_B_n8.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r2_bench_B_n8.json").read().strip().splitlines()[-1])
    print(d["value"], d["ms_per_step"], d["n_gpus"], d["clocks"], d["extra"].get("sp"))
except Exception as e:
    print("parse failed", e)
PY
