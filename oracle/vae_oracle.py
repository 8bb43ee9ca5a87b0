"""CPU oracle for the FLUX VAE decoder -- TEST INFRASTRUCTURE ONLY.

Restates ``models/modules/autoencoder.py``: AttnBlock :25-52, ResnetBlock :55-82, Downsample :85-95, Upsample :98-106,
Encoder :109-180, Decoder :183-259, DiagonalGaussian :262-275, AutoEncoder.encode/decode :302-309 as flat functions over a state dict with the
reference's (BFL) key names (``decoder.conv_in.weight``, ``decoder.mid.block_1.norm1.weight`` ...).
The pipeline itself calls diffusers' ``AutoencoderKL.decode`` (visualcloze.py:430), which is the
same architecture; diffusers is a third-party dependency absent from this image.

``dtype`` is the storage/compute dtype of activations and weights (bf16 in the pipeline,
``visualcloze.py:100``); GroupNorm statistics are taken in fp32 like ATen does.

Pinned against the reference module in ``oracle/gen_golden.py`` -> ``tests/golden/vae_*.pt``.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import torch
import torch.nn.functional as F


@dataclass
class VaeConfig:
    """AutoEncoderParams (autoencoder.py:8-18); defaults = FLUX VAE (models/util.py:154-164)."""

    ch: int = 128
    out_ch: int = 3
    ch_mult: list = field(default_factory=lambda: [1, 2, 4, 4])
    num_res_blocks: int = 2
    z_channels: int = 16
    scale_factor: float = 0.3611
    shift_factor: float = 0.1159
    groups: int = 32


def decoder_param_shapes(cfg: VaeConfig) -> dict[str, tuple]:
    shapes: dict[str, tuple] = {}

    def conv(name, cin, cout, k):
        shapes[name + ".weight"] = (cout, cin, k, k)
        shapes[name + ".bias"] = (cout,)

    def norm(name, c):
        shapes[name + ".weight"] = (c,)
        shapes[name + ".bias"] = (c,)

    def res(name, cin, cout):
        norm(name + ".norm1", cin); conv(name + ".conv1", cin, cout, 3)
        norm(name + ".norm2", cout); conv(name + ".conv2", cout, cout, 3)
        if cin != cout:
            conv(name + ".nin_shortcut", cin, cout, 1)

    nres = len(cfg.ch_mult)
    block_in = cfg.ch * cfg.ch_mult[-1]
    conv("decoder.conv_in", cfg.z_channels, block_in, 3)
    res("decoder.mid.block_1", block_in, block_in)
    norm("decoder.mid.attn_1.norm", block_in)
    for n in ("q", "k", "v", "proj_out"):
        conv(f"decoder.mid.attn_1.{n}", block_in, block_in, 1)
    res("decoder.mid.block_2", block_in, block_in)
    for lvl in reversed(range(nres)):
        block_out = cfg.ch * cfg.ch_mult[lvl]
        for b in range(cfg.num_res_blocks + 1):
            res(f"decoder.up.{lvl}.block.{b}", block_in, block_out)
            block_in = block_out
        if lvl != 0:
            conv(f"decoder.up.{lvl}.upsample.conv", block_in, block_in, 3)
    norm("decoder.norm_out", block_in)
    conv("decoder.conv_out", block_in, cfg.out_ch, 3)
    return shapes


def make_decoder_params(cfg: VaeConfig, seed: int = 0, dtype=torch.bfloat16) -> dict[str, torch.Tensor]:
    """Deterministic synthetic decoder weights (sorted-name order, one CPU generator)."""
    g = torch.Generator().manual_seed(seed)
    out = {}
    shapes = decoder_param_shapes(cfg)
    for name in sorted(shapes):
        shp = shapes[name]
        if len(shp) == 4:
            fan_in = shp[1] * shp[2] * shp[3]
            t = torch.randn(shp, generator=g) / math.sqrt(fan_in)
        elif ".norm" in name and name.endswith(".weight"):
            t = 1.0 + 0.1 * torch.randn(shp, generator=g)
        else:
            t = 0.02 * torch.randn(shp, generator=g)
        out[name] = t.to(dtype)
    return out


def _gn_swish(p, name, x, groups, swish=True):
    h = F.group_norm(x, groups, p[name + ".weight"], p[name + ".bias"], eps=1e-6)
    return h * torch.sigmoid(h) if swish else h


def _conv(p, name, x, pad):
    return F.conv2d(x, p[name + ".weight"], p[name + ".bias"], stride=1, padding=pad)


def _resnet(p, name, x, groups):
    h = _conv(p, name + ".conv1", _gn_swish(p, name + ".norm1", x, groups), 1)
    h = _conv(p, name + ".conv2", _gn_swish(p, name + ".norm2", h, groups), 1)
    if name + ".nin_shortcut.weight" in p:
        x = _conv(p, name + ".nin_shortcut", x, 0)
    return x + h


def _attn(p, name, x, groups):
    h = _gn_swish(p, name + ".norm", x, groups, swish=False)
    q, k, v = (_conv(p, f"{name}.{n}", h, 0) for n in ("q", "k", "v"))
    b, c, hh, ww = q.shape
    q, k, v = (t.reshape(b, c, hh * ww).transpose(1, 2) for t in (q, k, v))     # [b, P, c]
    s = torch.matmul(q.float(), k.float().transpose(1, 2)) * (c ** -0.5)
    a = torch.softmax(s, dim=-1)
    o = torch.matmul(a.to(v.dtype).float(), v.float()).to(x.dtype)
    o = o.transpose(1, 2).reshape(b, c, hh, ww)
    return x + _conv(p, name + ".proj_out", o, 0)


def decode(p: dict, cfg: VaeConfig, z: torch.Tensor, taps: dict | None = None) -> torch.Tensor:
    """AutoEncoder.decode: z [B, 16, h, w] -> image [B, 3, 8h, 8w] in roughly [-1, 1]."""
    g = cfg.groups
    z = z / cfg.scale_factor + cfg.shift_factor
    h = _conv(p, "decoder.conv_in", z, 1)
    h = _resnet(p, "decoder.mid.block_1", h, g)
    h = _attn(p, "decoder.mid.attn_1", h, g)
    h = _resnet(p, "decoder.mid.block_2", h, g)
    if taps is not None:
        taps["mid"] = h
    for lvl in reversed(range(len(cfg.ch_mult))):
        for b in range(cfg.num_res_blocks + 1):
            h = _resnet(p, f"decoder.up.{lvl}.block.{b}", h, g)
        if lvl != 0:
            h = F.interpolate(h, scale_factor=2.0, mode="nearest")
            h = _conv(p, f"decoder.up.{lvl}.upsample.conv", h, 1)
        if taps is not None:
            taps[f"up.{lvl}"] = h
    h = _gn_swish(p, "decoder.norm_out", h, g)
    return _conv(p, "decoder.conv_out", h, 1)


# ----------------------------------------------------------------------------------------------
# encoder ("next" row (f)-1 of SURVEY.md section 8)
# ----------------------------------------------------------------------------------------------
def encoder_param_shapes(cfg: VaeConfig, in_channels: int = 3) -> dict[str, tuple]:
    shapes: dict[str, tuple] = {}

    def conv(name, cin, cout, k):
        shapes[name + ".weight"] = (cout, cin, k, k)
        shapes[name + ".bias"] = (cout,)

    def norm(name, c):
        shapes[name + ".weight"] = (c,)
        shapes[name + ".bias"] = (c,)

    def res(name, cin, cout):
        norm(name + ".norm1", cin); conv(name + ".conv1", cin, cout, 3)
        norm(name + ".norm2", cout); conv(name + ".conv2", cout, cout, 3)
        if cin != cout:
            conv(name + ".nin_shortcut", cin, cout, 1)

    conv("encoder.conv_in", in_channels, cfg.ch, 3)
    in_mult = (1,) + tuple(cfg.ch_mult)
    block_in = cfg.ch
    for lvl in range(len(cfg.ch_mult)):
        block_in = cfg.ch * in_mult[lvl]
        block_out = cfg.ch * cfg.ch_mult[lvl]
        for b in range(cfg.num_res_blocks):
            res(f"encoder.down.{lvl}.block.{b}", block_in, block_out)
            block_in = block_out
        if lvl != len(cfg.ch_mult) - 1:
            conv(f"encoder.down.{lvl}.downsample.conv", block_in, block_in, 3)
    res("encoder.mid.block_1", block_in, block_in)
    norm("encoder.mid.attn_1.norm", block_in)
    for n in ("q", "k", "v", "proj_out"):
        conv(f"encoder.mid.attn_1.{n}", block_in, block_in, 1)
    res("encoder.mid.block_2", block_in, block_in)
    norm("encoder.norm_out", block_in)
    conv("encoder.conv_out", block_in, 2 * cfg.z_channels, 3)
    return shapes


def make_encoder_params(cfg: VaeConfig, seed: int = 0, dtype=torch.bfloat16) -> dict[str, torch.Tensor]:
    g = torch.Generator().manual_seed(seed)
    out = {}
    shapes = encoder_param_shapes(cfg)
    for name in sorted(shapes):
        shp = shapes[name]
        if len(shp) == 4:
            t = torch.randn(shp, generator=g) / math.sqrt(shp[1] * shp[2] * shp[3])
        elif ".norm" in name and name.endswith(".weight"):
            t = 1.0 + 0.1 * torch.randn(shp, generator=g)
        else:
            t = 0.02 * torch.randn(shp, generator=g)
        out[name] = t.to(dtype)
    return out


def encode_moments(p: dict, cfg: VaeConfig, x: torch.Tensor) -> torch.Tensor:
    """Encoder.forward: image [B, 3, H, W] -> moments [B, 2z, H/8, W/8] (mean | logvar)."""
    g = cfg.groups
    h = _conv(p, "encoder.conv_in", x, 1)
    for lvl in range(len(cfg.ch_mult)):
        for b in range(cfg.num_res_blocks):
            h = _resnet(p, f"encoder.down.{lvl}.block.{b}", h, g)
        if lvl != len(cfg.ch_mult) - 1:
            h = F.pad(h, (0, 1, 0, 1), mode="constant", value=0)                 # Downsample: asymmetric pad, stride 2
            h = F.conv2d(h, p[f"encoder.down.{lvl}.downsample.conv.weight"], p[f"encoder.down.{lvl}.downsample.conv.bias"], stride=2)
    h = _resnet(p, "encoder.mid.block_1", h, g)
    h = _attn(p, "encoder.mid.attn_1", h, g)
    h = _resnet(p, "encoder.mid.block_2", h, g)
    h = _gn_swish(p, "encoder.norm_out", h, g)
    return _conv(p, "encoder.conv_out", h, 1)


def encode(p: dict, cfg: VaeConfig, x: torch.Tensor, noise: torch.Tensor | None) -> torch.Tensor:
    """AutoEncoder.encode: z = scale * (mean + exp(0.5 logvar) * noise - shift)   (noise None -> mode)."""
    mean, logvar = torch.chunk(encode_moments(p, cfg, x), 2, dim=1)
    z = mean if noise is None else mean + torch.exp(0.5 * logvar) * noise
    return cfg.scale_factor * (z - cfg.shift_factor)
