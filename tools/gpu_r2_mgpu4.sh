#!/usr/bin/env bash
# round-2 session on FOUR GPUs: sequence-parallel tests at 2 and 4 ranks (4 ranks: persistent attention, 96-CTA per-pair grid),
# BASELINE config 5 (SDEdit 1024^2, 20 steps, batch 4 over 4 GPUs) and the contract's cfg-B line at N = 4 with extra.sp
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_sp_gpu.py -q -p no:cacheprovider --tb=short 2>&1 | tail -15 > gpurun_out/r2m4_sp_tests.log; tail -4 gpurun_out/r2m4_sp_tests.log
timeout 300 python -m pytest tests/test_fp8_gpu.py -q -p no:cacheprovider --tb=short -k "ln_modulate" 2>&1 | tail -5
run() { # workload, steps, tag
  timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29631 bench.py --gpus 4 --workload $1 --steps $2 --warmup 3 > "gpurun_out/r2_bench_$3.json" 2> "gpurun_out/r2_bench_$3.err"
  tail -c 500 "gpurun_out/r2_bench_$3.err"
  python - "gpurun_out/r2_bench_$3.json" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], d["value"], d["ms_per_step"], d["n_gpus"], d["config"]["workload"][:30], d["extra"].get("sp"))
except Exception as e:
    print("parse failed", sys.argv[1], e)
PY
}
run E 5 E_n4
run B 3 B_n4
