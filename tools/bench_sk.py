"""Stream-K tail on/off for the few-wave GEMMs of the sequence-parallel mode (M ~ 2000, N = 3072): run once per policy,
e.g.  VCB_STREAMK=0 python tools/bench_sk.py ; python tools/bench_sk.py   (CUDA events, L2 flushed between launches)."""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from visualcloze_b200 import ops  # noqa: E402

BF16 = torch.bfloat16
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
for M, N, K in ((1984, 3072, 12288), (1984, 3072, 15360), (1984, 3072, 3072), (992, 3072, 15360), (1984, 12288, 3072), (992, 3072, 12288), (496, 3072, 15360), (496, 3072, 12288), (992, 12288, 3072), (3968, 3072, 15360),
                (3968, 3072, 12288), (3968, 12288, 3072), (3968, 3072, 3072)):
    a = torch.randn(M, K, device="cuda").to(BF16)
    w = (torch.randn(N, K, device="cuda") / math.sqrt(K)).to(BF16)
    bias = torch.randn(N, device="cuda")
    gate = torch.randn(1, N, device="cuda").to(BF16)
    out = torch.randn(M, N, device="cuda").to(BF16)
    fn = lambda: ops.gemm(a, w, bias, out, epilogue=ops.EPI_GATE_RES, gate=gate, res=out)
    for _ in range(3):
        fn()
    ts = []
    for _ in range(20):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    print(f"VCB_STREAMK={os.environ.get('VCB_STREAMK', 'auto')} M={M} N={N} K={K}: median {ts[10] * 1e3:.1f} us, "
          f"{2.0 * M * N * K / ts[10] / 1e9:.0f} TFLOP/s", flush=True)
