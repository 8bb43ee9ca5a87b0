"""Generate tests/golden/*.pt by executing the UNMODIFIED reference modules -- build container only.

Run from the repo root:  ``python oracle/gen_golden.py``  (needs /root/reference; never runs on the
GPU box).  The reference cannot run on CPU as shipped, so exactly two symbols are
monkey-patched at run time (no reference file is edited or copied):

  * ``models.modules.layers.attention`` -> RoPE + ``F.scaled_dot_product_attention`` with a key
    mask, padded query rows zeroed (flash_attn_varlen_func is CUDA-only);
  * ``torch.cuda.device`` -> null context (layers.py:185,241 raise on CPU tensors).

Missing third-party modules are stubbed in ``sys.modules``: ``imwatermark`` (empty) and
``torchdiffeq`` (12-line fixed-grid Euler ``odeint`` with upstream semantics).

Weights come from ``oracle.flux_oracle.make_params`` (deterministic, seed-only), loaded into the
reference modules with ``load_state_dict(strict=True)`` -- which also pins the state-dict naming
contract.  Fixtures hold inputs + reference outputs only (small).
"""
from __future__ import annotations

import contextlib
import os
import sys
import types

import torch
import torch.nn.functional as F

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
sys.path.insert(0, REPO)
sys.path.insert(0, REF)

# ---- stubs for absent third-party modules ------------------------------------------------------
_wm = types.ModuleType("imwatermark")
_wm.WatermarkEncoder = type("WatermarkEncoder", (), {"set_watermark": lambda *a, **k: None})
sys.modules["imwatermark"] = _wm


def _odeint(func, y0, t, method="euler", atol=None, rtol=None):
    assert method == "euler"
    ys = [y0]
    y = y0
    for k in range(len(t) - 1):
        y = y + (t[k + 1] - t[k]) * func(t[k].to(y.abs().dtype), y)      # _PerturbFunc casts t to the state dtype
        ys.append(y)
    return torch.stack(ys, 0)


_td = types.ModuleType("torchdiffeq")
_td.odeint = _odeint
sys.modules["torchdiffeq"] = _td

import models.math as ref_math                      # noqa: E402
import models.modules.layers as ref_layers          # noqa: E402
from models.model import FluxLoraWrapper, FluxParams  # noqa: E402

from oracle import flux_oracle as fo                # noqa: E402
from oracle import vae_oracle as vo                 # noqa: E402


def _sdpa_attention(q, k, v, pe, attn_mask=None, drop_mask=None):
    q, k = ref_math.apply_rope(q, k, pe)
    m = None if attn_mask is None else (attn_mask[:, None, None, :] != 0)
    x = F.scaled_dot_product_attention(q, k, v, attn_mask=m)
    if attn_mask is not None:
        x = x * attn_mask[:, None, :, None].to(x.dtype)
    B, H, L, D = x.shape
    return x.permute(0, 2, 1, 3).reshape(B, L, H * D)


ref_layers.attention = _sdpa_attention
torch.cuda.device = lambda *a, **k: contextlib.nullcontext()

OUT = os.path.join(REPO, "tests", "golden")
os.makedirs(OUT, exist_ok=True)

SMALL = dict(in_channels=384, out_channels=64, vec_in_dim=32, context_in_dim=64, hidden_size=256,
             mlp_ratio=2.0, num_heads=2, depth=1, depth_single_blocks=1, axes_dim=[16, 56, 56],
             theta=10_000, qkv_bias=True, guidance_embed=True)
LORA_RANK = 16


def build_ref(dtype):
    cfg = fo.FluxConfig(**SMALL, lora_rank=LORA_RANK)
    model = FluxLoraWrapper(lora_rank=LORA_RANK, params=FluxParams(**SMALL)).eval()
    params = fo.make_params(cfg, seed=7)
    missing = model.load_state_dict({k: v.float() for k, v in params.items()}, strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    return cfg, model.to(dtype), params


def make_inputs(B, rows, seed, ragged=False):
    """rows: list of (h_tokens, w_tokens) per grid row; returns reference-shaped kwargs."""
    g = torch.Generator().manual_seed(seed)
    ids = []
    for j, (h, w) in enumerate(rows):
        t = torch.zeros(h, w, 3)
        t[..., 0] = j + 1
        t[..., 1] += torch.arange(h)[:, None]
        t[..., 2] += torch.arange(w)[None, :]
        ids.append(t.reshape(-1, 3))
    ids = torch.cat(ids, 0)
    Li, Lt = ids.shape[0], 24
    img = torch.randn(B, Li, 384, generator=g).bfloat16()
    txt = (0.5 * torch.randn(B, Lt, SMALL["context_in_dim"], generator=g)).bfloat16()
    y = torch.randn(B, SMALL["vec_in_dim"], generator=g).bfloat16()
    img_mask = torch.ones(B, Li, dtype=torch.int32)
    if ragged:
        img_mask[1, Li - 9:] = 0
    return dict(img=img, img_ids=ids[None].repeat(B, 1, 1), txt=txt, txt_ids=torch.zeros(B, Lt, 3),
                timesteps=torch.linspace(0.9, 0.3, B), y=y,
                txt_mask=torch.ones(B, Lt, dtype=torch.int32), img_mask=img_mask,
                guidance=torch.full((B,), 30.0, dtype=torch.bfloat16))


@torch.no_grad()
def gen_flux():
    for tag, B, rows, ragged in (("b1", 1, [(4, 6), (4, 6)], False), ("b2r", 2, [(4, 8)], True)):
        inp = make_inputs(B, rows, seed=11, ragged=ragged)
        # (1) fp32 reference, no autocast
        cfg, model, _ = build_ref(torch.float32)
        out32 = model(**{k: (v.float() if v.is_floating_point() else v) for k, v in inp.items()})
        # (2) reference under CPU bf16 autocast, bf16 weights (visualcloze.py:108,363 on CPU)
        cfg, model, _ = build_ref(torch.bfloat16)
        with torch.autocast("cpu", dtype=torch.bfloat16):
            out16 = model(**inp)
        torch.save({"inputs": inp, "out_fp32": out32, "out_cpu_bf16": out16,
                    "cfg": dict(SMALL, lora_rank=LORA_RANK), "param_seed": 7},
                   os.path.join(OUT, f"flux_small_{tag}.pt"))
        print(tag, out32.shape, out32.abs().mean().item(), out16.dtype)


@torch.no_grad()
def gen_sampler():
    from transport import Sampler, create_transport
    import models.sampling as ref_sampling
    cfg, model, _ = build_ref(torch.bfloat16)
    sampler = Sampler(create_transport("Linear", "velocity", do_shift=True))
    res = {}
    for tag, kw in (("shift4", dict(num_steps=4, do_shift=True, time_shifting_factor=1)),
                    ("sdedit5", dict(num_steps=5, do_shift=False, time_shifting_factor=1.0, strength=0.4))):
        inp = make_inputs(1, [(4, 6), (4, 6)], seed=13)
        g = torch.Generator().manual_seed(5)
        x = torch.randn(1, 48, 64, generator=g).bfloat16()
        cond = inp.pop("img")[..., 64:].contiguous()
        inp.pop("timesteps")
        seen_t = []
        fwd = model.forward

        def spy(x_, timesteps, **k):
            seen_t.append(timesteps.clone())
            return fwd(x_, timesteps=timesteps, **k)
        fn = sampler.sample_ode(sampling_method="euler", atol=1e-6, rtol=1e-3, reverse=False, **kw)
        mk = dict(inp, cond=cond)
        with torch.autocast("cpu", dtype=torch.bfloat16):
            traj = fn(x, spy, mk)
        assert "cond" in mk, "sampler must not mutate the caller's dict"
        res[tag] = dict(kwargs=kw, x=x, cond=cond, inputs=inp, traj=traj, flux_t=torch.stack(seen_t))
        print(tag, traj.shape, traj.dtype, [round(float(t), 5) for t in torch.stack(seen_t).flatten()])
    # schedule API goldens (models/sampling.py:300-328)
    sched = {}
    for n, L in ((29, 3456), (3, 576), (19, 4096), (29, 6912), (4, 256)):
        sched[(n, L)] = ref_sampling.get_schedule(n, L)
    sched["noshift"] = ref_sampling.get_schedule(10, 1024, shift=False)
    res["get_schedule"] = sched
    res["time_shift"] = ref_sampling.time_shift(1.0416667, 1.0, torch.linspace(1, 0, 7))
    res["lin_fn"] = [ref_sampling.get_lin_function()(v) for v in (256, 3456, 4096)]
    # token packing (sampling.py:37-118) with stub encoders
    t5 = lambda prompts: torch.arange(len(prompts) * 6 * 8, dtype=torch.float32).reshape(len(prompts), 6, 8)
    clip = lambda prompts: torch.ones(len(prompts), 5)
    g = torch.Generator().manual_seed(3)
    rows = [torch.randn(1, 16, 4, 12, generator=g), torch.randn(1, 16, 4, 12, generator=g)]
    rows_b = [torch.randn(1, 16, 4, 8, generator=g)]
    packed = ref_sampling.prepare_modified(t5, clip, [rows, rows_b], ["a", "b"], proportion_empty_prompts=0.0)
    res["prepare_modified"] = dict(rows=[rows, rows_b], out=packed)
    res["unpack"] = dict(x=packed["img"][:1, :12], out=ref_sampling.unpack(packed["img"][:1, :12], 32, 96))
    torch.save(res, os.path.join(OUT, "sampler.pt"))


@torch.no_grad()
def gen_vae():
    from models.modules.autoencoder import AutoEncoder, AutoEncoderParams
    small = dict(ch=64, out_ch=3, ch_mult=[1, 2], num_res_blocks=1, z_channels=16)   # level widths multiples of 64
    cfg = vo.VaeConfig(**small)
    ae = AutoEncoder(AutoEncoderParams(resolution=32, in_channels=3, scale_factor=0.3611,
                                       shift_factor=0.1159, **small)).eval()
    p = vo.make_decoder_params(cfg, seed=3, dtype=torch.float32)
    sd = ae.state_dict()
    dec_keys = [k for k in sd if k.startswith("decoder.")]
    assert sorted(dec_keys) == sorted(p), (set(dec_keys) ^ set(p))
    ae.load_state_dict({**sd, **p}, strict=True)
    g = torch.Generator().manual_seed(9)
    z = torch.randn(1, 16, 12, 20, generator=g)
    out = ae.decode(z)
    # encoder: moments of the reference Encoder on a seeded image, same small geometry
    pe = vo.make_encoder_params(cfg, seed=4, dtype=torch.float32)
    enc_keys = [k for k in sd if k.startswith("encoder.")]
    assert sorted(enc_keys) == sorted(pe), (set(enc_keys) ^ set(pe))
    ae.load_state_dict({**ae.state_dict(), **pe}, strict=True)
    img = torch.randn(1, 3, 32, 48, generator=g).clamp(-1, 1)
    mom = ae.encoder(img)
    torch.save({"cfg": small, "param_seed": 3, "z": z, "out_fp32": out, "enc_param_seed": 4, "img": img, "moments_fp32": mom},
               os.path.join(OUT, "vae_small.pt"))
    print("vae", out.shape, out.abs().mean().item(), "moments", mom.shape, mom.abs().mean().item())


if __name__ == "__main__":
    torch.manual_seed(0)
    gen_flux()
    gen_sampler()
    gen_vae()
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))
