#!/usr/bin/env bash
# A/B of the epilogue restructuring: the in-tree library vs tools/_probe/libvcb200_new.so on the six block GEMM shapes
mkdir -p gpurun_out
python tools/fp8_gemm_probe.py > gpurun_out/r2l_probe_old.log 2>&1; tail -7 gpurun_out/r2l_probe_old.log | head -6
python tools/fp8_gemm_probe.py tools/_probe/libvcb200_new.so > gpurun_out/r2l_probe_new.log 2>&1; tail -7 gpurun_out/r2l_probe_new.log | head -6
