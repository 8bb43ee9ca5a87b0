#!/usr/bin/env bash
# closing regression with ABI 6 (text encoders added): smoke(), the whole -m gpu suite, the headline bench line
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 1500 python -m pytest tests -x -q -m gpu -p no:cacheprovider --tb=short -rf 2>&1 | grep -v Warning | tail -12 > gpurun_out/r2u_pytest.log
tail -4 gpurun_out/r2u_pytest.log
timeout 900 python bench.py > gpurun_out/r2u_bench.json 2> gpurun_out/r2u_bench.err; tail -c 300 gpurun_out/r2u_bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2u_bench.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step")}, d["e2e"]["value"], d["clocks"]["sm_mhz"], d["roofline"]["frac"], d["roofline_attention"]["frac"])
PY
