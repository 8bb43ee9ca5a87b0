"""Launch ONE hot kernel of the path at its cfg-B shape a few times, for `ncu --set full -k regex:... -c 1` captures.

  python tools/ncu_targets.py ln | ln_stats | attn_pair | attn_pair_exact | attn_persistent | gemm_linear1 | gemm_linear2 | gemm_fp8_linear1 |
                              gemm_fp8_linear2 | quantize_cat | conv512 | conv256 | conv128
"""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from visualcloze_b200 import _lib, ops  # noqa: E402

BF16 = torch.bfloat16
H, MLP, Li, Lt = 3072, 12288, 3456, 512
L = Li + Lt
what = sys.argv[1]
g = torch.Generator(device="cuda").manual_seed(0)
rn = lambda *s: torch.randn(*s, generator=g, device="cuda")
REP = 3

if what == "ln":
    x, y = rn(L, H).to(BF16), torch.empty(L, H, dtype=BF16, device="cuda")
    mods = [(0.2 * rn(1, 6 * H)).to(BF16) for _ in range(2)]
    for _ in range(REP):   # the img + txt LayerNorms of a double block in one launch
        ops.ln_modulate_grouped(x, y, [(Lt, Li, Li, mods[0][:, :H], mods[0][:, H:2 * H]), (0, Lt, Lt, mods[1][:, :H], mods[1][:, H:2 * H])], H, L, 6 * H)
elif what == "ln_stats":
    # the LayerNorm that follows a residual GEMM: statistics come from the GEMM's epilogue (row_stats), the row is streamed once
    x, y = rn(L, H).to(BF16), torch.empty(L, H, dtype=BF16, device="cuda")
    xf = x.float().reshape(L, H // 64, 64)
    stats = torch.stack((xf.sum(-1), (xf * xf).sum(-1)), dim=-1).contiguous()
    mod = (0.2 * rn(1, 3 * H)).to(BF16)
    for _ in range(REP):
        ops.ln_modulate_stats(x, mod[:, :H], mod[:, H:2 * H], y, stats, rows_per_batch=L, mod_stride=3 * H)
elif what.startswith("attn"):
    qkv = rn(L, 3, 24, 128)
    for i, a in ((0, 1.2), (1, 1.1)):
        qkv[:, i] = a * qkv[:, i] / qkv[:, i].pow(2).mean(-1, keepdim=True).sqrt()
    qkv = qkv.reshape(L, 3 * H).to(BF16)
    out = torch.empty(L, H, dtype=BF16, device="cuda")
    bound = 0.0 if what.endswith("exact") else 1.2 * 1.1 * math.sqrt(128.0) * math.log2(math.e) * 1.03
    sched = 2 if "persistent" in what else 1
    for _ in range(REP):
        ops.attention(qkv, 1, L, 24, out, q_col=0, k_col=H, v_col=2 * H, score_bound_log2=bound, schedule=sched)
elif what.startswith("gemm"):
    rope = torch.zeros(64, L, 2, device="cuda"); rope[..., 0] = 1.0
    qs = torch.ones(128, dtype=BF16, device="cuda")
    x = rn(L, H).to(BF16)
    qkv = torch.empty(L, 3 * H, dtype=BF16, device="cuda")
    cat = rn(L, H + MLP).to(BF16)
    if what == "gemm_fp8_linear2":
        # fp8 level 2: linear2 (K = 15360, gated residual) on the e4m3 copy of cat that quantize_rows_e4m3_kernel produced
        w32, b, gate = rn(H, H + MLP) / math.sqrt(H + MLP), rn(H), (0.3 * rn(1, H)).to(BF16)
        sw = w32.abs().amax(1) / 448.0
        w8 = (w32 / sw[:, None]).to(torch.float8_e4m3fn)
        c8, sa = torch.empty(L, H + MLP, dtype=torch.float8_e4m3fn, device="cuda"), torch.empty(L, device="cuda")
        ops.quantize_rows_e4m3(cat, c8, sa)
        stats = torch.zeros(L, H // 64, 2, device="cuda")
        for _ in range(REP):
            ops.gemm(c8, w8, b, x, epilogue=ops.EPI_GATE_RES, gate=gate, res=x, a_scale=sa, w_scale=sw, row_stats=stats)
    elif what == "gemm_linear2":
        w, b, gate = (rn(H, H + MLP) / math.sqrt(H + MLP)).to(BF16), rn(H), (0.3 * rn(1, H)).to(BF16)
        for _ in range(REP):
            ops.gemm(cat, w, b, x, epilogue=ops.EPI_GATE_RES, gate=gate, res=x)
    else:
        N = 3 * H + MLP
        w32 = rn(N, H) / math.sqrt(H)
        b = rn(N)
        kw = dict(epilogue=ops.EPI_LINEAR1, hidden=H, q_scale=qs, k_scale=qs, rope=rope, out2=cat, out2_col_offset=H)
        if what == "gemm_fp8_linear1":
            sw = w32.abs().amax(1) / 448.0
            w8 = (w32 / sw[:, None]).to(torch.float8_e4m3fn)
            sa = x.float().abs().amax(1) / 448.0
            a8 = (x.float() / sa[:, None]).to(torch.float8_e4m3fn)
            for _ in range(REP):
                ops.gemm(a8, w8, b, qkv, a_scale=sa, w_scale=sw, **kw)
        else:
            w = w32.to(BF16)
            for _ in range(REP):
                ops.gemm(x, w, b, qkv, **kw)
elif what == "quantize_cat":
    cat = rn(L, H + MLP).to(BF16)
    c8, sa = torch.empty(L, H + MLP, dtype=torch.float8_e4m3fn, device="cuda"), torch.empty(L, device="cuda")
    for _ in range(REP):
        ops.quantize_rows_e4m3(cat, c8, sa)
elif what.startswith("conv"):
    C = int(what[4:])
    Hh, Ww = {512: (96, 288), 256: (192, 576), 128: (384, 1152)}[C]     # decoder levels of one cfg-B grid row
    xg = rn(1, Hh, Ww, C).to(BF16)
    wg = (rn(C, 9 * C) / math.sqrt(9 * C)).to(BF16)
    b = rn(C)
    out = torch.empty(1, Hh, Ww, C, dtype=BF16, device="cuda")
    for _ in range(REP):
        _lib.check(_lib.lib().vcb_conv3x3_nhwc(xg.data_ptr(), wg.data_ptr(), b.data_ptr(), None, out.data_ptr(), 1, Hh, Ww, C, C, 1, None), "conv")
else:
    raise SystemExit(f"unknown target {what}")
torch.cuda.synchronize()
print("ran", what)
