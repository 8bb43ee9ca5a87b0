#!/usr/bin/env bash
# Round-end evidence on one B200, from the repo root: full GPU test suite, smoke(), the default bench line, the ncu launch
# list and one full-set capture each of the dominant GEMMs and of the attention kernel.  Everything lands in gpurun_out/.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/final_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/final_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/final_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/final_smoke.log
timeout 600 python bench.py > gpurun_out/bench_r01_n1.json 2> gpurun_out/bench_r01_n1.err; echo "bench rc=$?"; cut -c1-400 gpurun_out/bench_r01_n1.json
KREGEX='regex:gemm_bf16_tcgen05|attn_fwd|ln_modulate|euler_update|silu_kernel|add3|rope_table|timestep_embedding|copy_cols|vae_|gn_|conv'
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KREGEX" -c 1200 --csv --log-file gpurun_out/launches.csv \
    python tools/time_full.py 3 > gpurun_out/ncu_launches.log 2>&1; echo "launch list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_tcgen05 -s 200 -c 4 -f -o gpurun_out/prof_gemm \
    python tools/time_full.py 2 > gpurun_out/ncu_gemm.log 2>&1; echo "ncu gemm rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_fwd -s 3 -c 2 -f -o gpurun_out/prof_attn \
    python tools/time_full.py 2 > gpurun_out/ncu_attn.log 2>&1; echo "ncu attn rc=$?"
ls -la gpurun_out | head -40
