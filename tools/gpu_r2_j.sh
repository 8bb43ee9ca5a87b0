#!/usr/bin/env bash
# after packing oracle/_ref as an archive: the tests that use it, the reference arm with the thread-count search, one default bench line
mkdir -p gpurun_out
nproc; cat /sys/fs/cgroup/cpu.max 2>/dev/null; python -c "import os; print('affinity', len(os.sched_getaffinity(0)), 'cpu_count', os.cpu_count())"
timeout 1200 python -m pytest tests/test_fullsize_gpu.py tests/test_pipeline_gpu.py -q -m gpu -p no:cacheprovider --tb=short -rf -s 2>&1 | grep -v Warning | tail -40 > gpurun_out/r2j_pytest.log
tail -25 gpurun_out/r2j_pytest.log
timeout 900 python bench.py --impl reference > gpurun_out/r2j_bench_reference.json 2> gpurun_out/r2j_bench_reference.err; tail -c 1200 gpurun_out/r2j_bench_reference.json
timeout 900 python bench.py > gpurun_out/r2j_bench.json 2> gpurun_out/r2j_bench.err; tail -c 300 gpurun_out/r2j_bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2j_bench.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step")}, d["clocks"]["sm_mhz"], d["cpu_baseline"])
PY
