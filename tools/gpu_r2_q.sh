#!/usr/bin/env bash
# closing check on EIGHT GPUs with the final library: the contract's cfg-B line exactly as the driver launches it (replicas + extra.sp:
# one image over 8 GPUs through the sequence-parallel kernels), then the fp8_all line on the same box
mkdir -p gpurun_out
run() { # tag, extra args
  timeout 800 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29641 bench.py --gpus 8 --steps 5 --warmup 3 $2 > gpurun_out/r2q_bench_n8$1.json 2> gpurun_out/r2q_bench_n8$1.err
  tail -c 300 gpurun_out/r2q_bench_n8$1.err
  python - "gpurun_out/r2q_bench_n8$1.json" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], d["value"], d["ms_per_step"], d["n_gpus"], d["e2e"]["value"], d["clocks"], d["extra"].get("sp"))
except Exception as e:
    print("parse failed", e)
PY
}
run "" ""
run "_fp8_all" "--precision fp8_all"
