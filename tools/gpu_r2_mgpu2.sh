#!/usr/bin/env bash
# round-2 GPU session on TWO GPUs: sequence-parallel tests (staged TMA attention epilogue, then the persistent attention kernel),
# the contract's bench line at N = 2 (replicas + extra.sp), and the same with the persistent attention in the SP part
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader
timeout 900 python -m pytest tests/test_sp_gpu.py -q -p no:cacheprovider 2>&1 | tail -15 > gpurun_out/r2m_sp_tests.log; tail -5 gpurun_out/r2m_sp_tests.log
VCB_SP_ATTN_PERSIST=1 timeout 900 python -m pytest tests/test_sp_gpu.py -q -p no:cacheprovider 2>&1 | tail -15 > gpurun_out/r2m_sp_tests_persist.log; tail -5 gpurun_out/r2m_sp_tests_persist.log
run() { # tag, extra env
  env $2 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 3 --warmup 3 > "gpurun_out/r2m_bench_n2$1.json" 2> "gpurun_out/r2m_bench_n2$1.err"
  tail -c 600 "gpurun_out/r2m_bench_n2$1.err"
  python - "gpurun_out/r2m_bench_n2$1.json" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], d["value"], d["ms_per_step"], d["n_gpus"], d["extra"].get("sp"))
except Exception as e:
    print("parse failed", sys.argv[1], e)
PY
}
run "" "VCB_X=0"
run "_sp_persist" "VCB_SP_ATTN_PERSIST=1"
