"""Bounded attention kernel at L = 3968 / 7424 (CUDA events, L2 flushed)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from visualcloze_b200 import ops
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
for L in (3968, 7424):
    H = 24 * 128
    qkv = torch.randn(L, 3 * H, device="cuda").bfloat16()
    out = torch.empty(L, H, dtype=torch.bfloat16, device="cuda")
    fn = lambda: ops.attention(qkv, 1, L, 24, out, q_col=0, k_col=H, v_col=2 * H, score_bound_log2=30.0)
    for _ in range(3): fn()
    ts = []
    for _ in range(15):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    print(f"L={L}: {ts[7] * 1e3:.1f} us, {4.0 * L * L * H / ts[7] / 1e9:.0f} TFLOP/s", flush=True)
