#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 600 python tools/bench_text_encoders.py 2>&1 | tail -3
