"""Sustained per-shape timing of the six block GEMMs of cfg B (real fused epilogues), bf16 and e4m3 operands, for A/B-ing a
kernel build:  python tools/fp8_gemm_probe.py [path/to/libvcb200.so]   (default: the in-tree library).
Each shape runs back to back for ~0.4 s (the chip sits at its power cap like in the loop); prints us and TFLOP/s."""
import json
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from visualcloze_b200 import _lib  # noqa: E402

if len(sys.argv) > 1:
    _lib.LIB_PATH = os.path.abspath(sys.argv[1])
from visualcloze_b200 import ops  # noqa: E402

BF16, F8 = torch.bfloat16, torch.float8_e4m3fn
H, MLP, L = 3072, 12288, 3968
g = torch.Generator(device="cuda").manual_seed(0)
rn = lambda *s: torch.randn(*s, generator=g, device="cuda")


def sustained(fn, seconds=0.4):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); fn(); b.record(); torch.cuda.synchronize()
    n = max(10, int(seconds * 1e3 / max(a.elapsed_time(b), 1e-3)))
    a.record()
    for _ in range(n):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3


def quant(x):
    s = x.float().abs().amax(1).clamp_min(1e-12) / 448.0
    return (x.float() / s[:, None]).to(F8), s


rope = torch.zeros(64, L, 2, device="cuda"); rope[..., 0] = 1.0
qs = torch.ones(128, dtype=BF16, device="cuda")
x = rn(L, H).to(BF16)
qkv = torch.empty(L, 3 * H, dtype=BF16, device="cuda")
cat = rn(L, H + MLP).to(BF16)
gate = (0.3 * rn(1, H)).to(BF16)
stats = torch.zeros(L, H // 64, 2, device="cuda")
out = {}
cases = [("qkv", 3 * H, H, dict(epilogue=ops.EPI_QKV, hidden=H, q_scale=qs, k_scale=qs, rope=rope), "qkv"),
         ("proj", H, H, dict(epilogue=ops.EPI_GATE_RES, gate=gate, res=x, row_stats=stats), "x"),
         ("mlp0_gelu", MLP, H, dict(epilogue=ops.EPI_BIAS_GELU, out_col_offset=H), "cat"),
         ("mlp2", H, MLP, dict(epilogue=ops.EPI_GATE_RES, gate=gate, res=x, row_stats=stats), "x"),
         ("linear1", 3 * H + MLP, H, dict(epilogue=ops.EPI_LINEAR1, hidden=H, q_scale=qs, k_scale=qs, rope=rope, out2=cat, out2_col_offset=H), "qkv"),
         ("linear2", H, H + MLP, dict(epilogue=ops.EPI_GATE_RES, gate=gate, res=x, row_stats=stats), "x")]
for name, N, K, kw, dst in cases:
    a = rn(L, K).to(BF16)
    w = (rn(N, K) / math.sqrt(K)).to(BF16)
    b = rn(N)
    o = {"qkv": qkv, "x": x, "cat": cat}[dst]
    a8, sa = quant(a)
    w8, sw = quant(w)
    fl = 2.0 * L * N * K
    t16 = sustained(lambda: ops.gemm(a, w, b, o, **kw))
    t8 = sustained(lambda: ops.gemm(a8, w8, b, o, a_scale=sa, w_scale=sw, **kw))
    out[name] = dict(N=N, K=K, bf16_us=round(t16, 1), bf16_tflops=round(fl / t16 / 1e6), fp8_us=round(t8, 1), fp8_tflops=round(fl / t8 / 1e6))
    print(name, out[name], flush=True)
print(json.dumps(out))
