#!/usr/bin/env bash
# closing check on TWO GPUs with the final library: sequence-parallel tests, the contract's N = 2 line exactly as the driver launches
# it (default --steps / --warmup come from the driver; 5 / 3 here), and the reference arm under torchrun (rank 0 prints, rank 1 exits)
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader
timeout 900 python -m pytest tests/test_sp_gpu.py -q -p no:cacheprovider 2>&1 | tail -15 > gpurun_out/r2o_sp_tests.log; tail -4 gpurun_out/r2o_sp_tests.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2o_bench_n2.json 2> gpurun_out/r2o_bench_n2.err
tail -c 400 gpurun_out/r2o_bench_n2.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r2o_bench_n2.json").read().strip().splitlines()[-1])
    print(d["value"], d["ms_per_step"], d["n_gpus"], d["e2e"], d["clocks"], d["extra"].get("sp"))
except Exception as e:
    print("parse failed", e)
PY
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --impl reference --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2o_bench_reference_n2.json 2> gpurun_out/r2o_bench_reference_n2.err
echo "reference arm rc=$?"; tail -c 500 gpurun_out/r2o_bench_reference_n2.json
