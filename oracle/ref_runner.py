"""Run the UNMODIFIED reference modules from ``oracle/_ref`` (see ``oracle/build_ref.py``) -- TEST INFRASTRUCTURE ONLY.

This is the authoritative oracle of SURVEY.md 8c: ``models/model.py::FluxLoraWrapper`` under ``torch.autocast("cuda", bf16)``
with flash-attn (``visualcloze.py:363``, ``models/math.py:85-95``), the reference ``transport`` sampler on the torchdiffeq
stand-in, and ``models/modules/autoencoder.py::AutoEncoder``.  Only ``tests/`` and ``bench.py --impl reference`` /
``cpu_baseline`` may import this module; nothing under ``visualcloze_b200/`` does.

On CPU (bench reference arm, golden generation) the reference needs the two run-time patches of SURVEY.md 8c -- no file of
``oracle/_ref`` is edited: ``models.modules.layers.attention`` -> RoPE + SDPA with a key mask (flash-attn is CUDA-only), and
``torch.cuda.device`` -> null context (layers.py:185,241 raise on CPU tensors).
"""
from __future__ import annotations

import contextlib
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "_ref")
ARCHIVE = os.path.join(REF, "reference_modules.zip")      # the unmodified models/ and transport/ packages, imported by zipimport


def available() -> bool:
    return os.path.exists(os.path.join(REF, "MANIFEST.json")) and os.path.exists(ARCHIVE)


def _import():
    if not available():
        raise RuntimeError("oracle/_ref is missing: run `python oracle/build_ref.py` in the build container (needs /root/reference)")
    from oracle import build_ref
    build_ref.verify()
    for entry in (REF, ARCHIVE):                       # stand-ins (torchdiffeq, imwatermark) as files, the reference from the archive
        if entry not in sys.path:
            sys.path.insert(0, entry)
    import models.model as ref_model
    import models.modules.autoencoder as ref_ae
    import models.modules.layers as ref_layers
    import transport as ref_transport
    return ref_model, ref_layers, ref_ae, ref_transport


def apply_cpu_patches():
    """The two monkey-patches that let the reference blocks run on CPU tensors (see module docstring)."""
    import torch.nn.functional as F
    _, ref_layers, _, _ = _import()
    import models.math as ref_math

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


def build_flux(params_kwargs: dict, lora_rank: int, state_dict: dict, lora_scale: float = 1.0):
    """Reference ``FluxLoraWrapper`` whose parameters ARE the given tensors (meta construction + ``assign=True``: no second copy of
    a 24 GB model).  ``strict=True`` also pins the state-dict naming contract."""
    ref_model, _, _, _ = _import()
    with torch.device("meta"):
        m = ref_model.FluxLoraWrapper(lora_rank=lora_rank, lora_scale=float(lora_scale), params=ref_model.FluxParams(**params_kwargs))
    res = m.load_state_dict(state_dict, strict=True, assign=True)
    assert not res.missing_keys and not res.unexpected_keys
    return m.eval().requires_grad_(False)


@torch.no_grad()
def flux_forward(model, **inputs):
    """``Flux.forward`` exactly as the pipeline calls it: inside ``torch.autocast("cuda", bf16)`` (visualcloze.py:363)."""
    dev = next(model.parameters()).device
    with torch.autocast(dev.type, dtype=torch.bfloat16):
        return model(**inputs)


@torch.no_grad()
def sample_ode(model, x, model_kwargs: dict, *, num_steps: int, do_shift: bool = True, time_shifting_factor=1, strength=None):
    """Reference ``Sampler(create_transport("Linear","velocity")).sample_ode(...)`` (visualcloze.py:118-131, 226-234) driving the
    reference model's bound ``forward``; returns the stacked trajectory [num_steps, B, Li, 64]."""
    _, _, _, ref_transport = _import()
    sampler = ref_transport.Sampler(ref_transport.create_transport("Linear", "velocity", do_shift=True))
    kw = dict(sampling_method="euler", num_steps=num_steps, atol=1e-6, rtol=1e-3, reverse=False, do_shift=do_shift,
              time_shifting_factor=time_shifting_factor)
    if strength is not None:
        kw["strength"] = strength
    fn = sampler.sample_ode(**kw)
    dev = next(model.parameters()).device
    with torch.autocast(dev.type, dtype=torch.bfloat16):
        return fn(x, model.forward, dict(model_kwargs))


def build_autoencoder(ae_kwargs: dict, state_dict: dict, device, dtype=torch.bfloat16):
    """In-repo ``AutoEncoder`` (autoencoder.py:262-312; same architecture as the diffusers FLUX VAE the pipeline loads)."""
    _, _, ref_ae, _ = _import()
    ae = ref_ae.AutoEncoder(ref_ae.AutoEncoderParams(**ae_kwargs))
    own = ae.state_dict()
    res = ae.load_state_dict({**own, **state_dict}, strict=True)
    assert not res.missing_keys
    return ae.to(device=device, dtype=dtype).eval().requires_grad_(False)


def flash_attention(q, k, v):
    """flash-attn exactly as models/math.py:85-95 calls it for an unpadded batch: q, k, v [B, L, H, D] bf16 -> [B, L, H, D]."""
    from flash_attn import flash_attn_varlen_func
    B, L, H, D = q.shape
    cu = torch.arange(0, (B + 1) * L, L, dtype=torch.int32, device=q.device)
    o = flash_attn_varlen_func(q.reshape(B * L, H, D), k.reshape(B * L, H, D), v.reshape(B * L, H, D), cu, cu, L, L,
                               dropout_p=0.0, causal=False)
    return o.reshape(B, L, H, D)
