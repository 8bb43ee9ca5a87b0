#!/usr/bin/env bash
# round-2 final validation on one GPU: smoke(), the whole -m gpu suite, the contract's bench line (both arms) and the fp8 line
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider --tb=short -rf 2>&1 | tail -40 > gpurun_out/r2z_pytest.log
tail -6 gpurun_out/r2z_pytest.log
show() { python - "$1" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], {k: d[k] for k in ("value", "ms_per_step", "dtype")}, "e2e", d["e2e"]["value"], d["clocks"]["sm_mhz"])
    for k in ("roofline", "roofline_attention", "roofline_ln_modulate", "roofline_vae"):
        print(" ", k, round(d[k]["achieved"], 1), round(d[k]["frac"], 3))
    print(" ", d["kernel_time_share"]); print(" ", d["extra"]); print(" ", d.get("cpu_baseline"))
except Exception as e:
    print("bench parse failed", sys.argv[1], e)
PY
}
timeout 900 python bench.py > gpurun_out/r2z_bench.json 2> gpurun_out/r2z_bench.err; tail -c 400 gpurun_out/r2z_bench.err; show gpurun_out/r2z_bench.json
timeout 900 python bench.py --precision fp8 --no-cpu-baseline > gpurun_out/r2z_bench_fp8.json 2> gpurun_out/r2z_bench_fp8.err; show gpurun_out/r2z_bench_fp8.json
timeout 600 python bench.py --impl reference > gpurun_out/r2z_bench_reference.json 2> gpurun_out/r2z_bench_reference.err; tail -c 300 gpurun_out/r2z_bench_reference.json
