"""FLUX VAE decoder on libvcb200 with the reference's decoder state-dict names and ``decode`` semantics.

Mirror of ``models/modules/autoencoder.py`` (``AutoEncoderParams`` :8-18, ``Decoder`` :183-259,
``AutoEncoder.decode`` :307-309) -- the same architecture the pipeline reaches through diffusers'
``AutoencoderKL.decode`` (``visualcloze.py:430``).  Parameters are registered under the reference's
BFL key names (``decoder.conv_in.weight`` ... ``decoder.up.{lvl}.block.{b}.conv1.weight``), so BFL ``ae.safetensors``
loads with ``load_state_dict(strict=False)`` (encoder keys are ignored; VAE *encode* is a "next" row).
"""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass, field

import torch
from torch import Tensor, nn

from . import _lib
from ._lib import ConvW, GnW, ResblockW, VaeConfigC, VaeEncWeightsC, VaeWeightsC, check

BF16 = torch.bfloat16


@dataclass
class AutoEncoderParams:
    resolution: int = 256
    in_channels: int = 3
    ch: int = 128
    out_ch: int = 3
    ch_mult: list = field(default_factory=lambda: [1, 2, 4, 4])
    num_res_blocks: int = 2
    z_channels: int = 16
    scale_factor: float = 0.3611
    shift_factor: float = 0.1159


def decoder_param_shapes(p: AutoEncoderParams) -> dict[str, tuple]:
    sh: dict[str, tuple] = {}

    def conv(n, ci, co, k):
        sh[n + ".weight"], sh[n + ".bias"] = (co, ci, k, k), (co,)

    def norm(n, c):
        sh[n + ".weight"], sh[n + ".bias"] = (c,), (c,)

    def res(n, ci, co):
        norm(n + ".norm1", ci); conv(n + ".conv1", ci, co, 3); norm(n + ".norm2", co); conv(n + ".conv2", co, co, 3)
        if ci != co:
            conv(n + ".nin_shortcut", ci, co, 1)

    bi = p.ch * p.ch_mult[-1]
    conv("decoder.conv_in", p.z_channels, bi, 3)
    res("decoder.mid.block_1", bi, bi)
    norm("decoder.mid.attn_1.norm", bi)
    for n in ("q", "k", "v", "proj_out"):
        conv(f"decoder.mid.attn_1.{n}", bi, bi, 1)
    res("decoder.mid.block_2", bi, bi)
    for lvl in reversed(range(len(p.ch_mult))):
        bo = p.ch * p.ch_mult[lvl]
        for b in range(p.num_res_blocks + 1):
            res(f"decoder.up.{lvl}.block.{b}", bi, bo)
            bi = bo
        if lvl != 0:
            conv(f"decoder.up.{lvl}.upsample.conv", bi, bi, 3)
    norm("decoder.norm_out", bi)
    conv("decoder.conv_out", bi, p.out_ch, 3)
    return sh


def diffusers_vae_to_bfl(sd: dict) -> dict:
    """Rename diffusers ``AutoencoderKL`` keys (FLUX.1-dev ``vae/``; what visualcloze.py:100 loads) to the BFL / in-repo
    ``AutoEncoder`` names (models/modules/autoencoder.py).  Pure renaming plus the attention projections' Linear [C, C] ->
    1x1 conv [C, C, 1, 1] view; the up path is numbered in the opposite direction (``up_blocks.i`` == ``up.(n-1-i)``)."""
    import re
    n_up = 1 + max([int(m.group(1)) for k in sd for m in [re.match(r"decoder\.up_blocks\.(\d+)\.", k)] if m] or [3])
    res_names = {"conv_shortcut": "nin_shortcut"}
    attn_names = {"group_norm": "norm", "to_q": "q", "to_k": "k", "to_v": "v", "to_out.0": "proj_out"}
    out = {}
    for k, v in sd.items():
        m = re.match(r"(encoder|decoder)\.(.*)", k)
        if not m:
            continue                                     # quant_conv / post_quant_conv: absent in the FLUX VAE
        half, rest = m.groups()
        nk = None
        if (mm := re.match(r"down_blocks\.(\d+)\.resnets\.(\d+)\.(\w+)\.(weight|bias)", rest)):
            i, j, what, wb = mm.groups()
            nk = f"down.{i}.block.{j}.{res_names.get(what, what)}.{wb}"
        elif (mm := re.match(r"down_blocks\.(\d+)\.downsamplers\.0\.conv\.(weight|bias)", rest)):
            nk = f"down.{mm.group(1)}.downsample.conv.{mm.group(2)}"
        elif (mm := re.match(r"up_blocks\.(\d+)\.resnets\.(\d+)\.(\w+)\.(weight|bias)", rest)):
            i, j, what, wb = mm.groups()
            nk = f"up.{n_up - 1 - int(i)}.block.{j}.{res_names.get(what, what)}.{wb}"
        elif (mm := re.match(r"up_blocks\.(\d+)\.upsamplers\.0\.conv\.(weight|bias)", rest)):
            nk = f"up.{n_up - 1 - int(mm.group(1))}.upsample.conv.{mm.group(2)}"
        elif (mm := re.match(r"mid_block\.resnets\.(\d+)\.(\w+)\.(weight|bias)", rest)):
            j, what, wb = mm.groups()
            nk = f"mid.block_{int(j) + 1}.{res_names.get(what, what)}.{wb}"
        elif (mm := re.match(r"mid_block\.attentions\.0\.(group_norm|to_q|to_k|to_v|to_out\.0)\.(weight|bias)", rest)):
            what, wb = mm.groups()
            nk = f"mid.attn_1.{attn_names[what]}.{wb}"
            if wb == "weight" and what != "group_norm" and v.dim() == 2:
                v = v[:, :, None, None]
        elif (mm := re.match(r"conv_norm_out\.(weight|bias)", rest)):
            nk = f"norm_out.{mm.group(1)}"
        elif re.match(r"(conv_in|conv_out)\.(weight|bias)", rest):
            nk = rest
        if nk is not None:
            out[f"{half}.{nk}"] = v
    return out


def _attach(root: nn.Module, dotted: str, param: nn.Parameter) -> None:
    node = root
    parts = dotted.split(".")
    for name in parts[:-1]:
        if name not in node._modules:
            node.add_module(name, nn.Module())
        node = node._modules[name]
    node.register_parameter(parts[-1], param)


class AutoEncoderDecoder(nn.Module):
    def __init__(self, params: AutoEncoderParams | None = None, device=None, dtype=BF16):
        super().__init__()
        self.params = params or AutoEncoderParams()
        self.scale_factor, self.shift_factor = self.params.scale_factor, self.params.shift_factor
        for name, shp in decoder_param_shapes(self.params).items():
            _attach(self, name, nn.Parameter(torch.empty(shp, device=device, dtype=dtype), requires_grad=False))
        self._h = None
        self._key = None
        self._ws = None
        self._initialized: set[str] = set()

    # ---- weights: parameters are allocated with torch.empty; packing never-loaded memory must fail loudly ---------------
    def load_state_dict(self, state_dict, strict: bool = True, assign: bool = False):
        """Accepts the BFL ``ae.safetensors`` names (== models/modules/autoencoder.py) restricted to this half's prefix, or
        diffusers ``AutoencoderKL`` names (what the reference pipeline loads, visualcloze.py:100) -- converted on the fly."""
        prefix = "encoder." if isinstance(self, AutoEncoderEncoder) else "decoder."
        sd = dict(state_dict)
        if any(".up_blocks." in k or ".down_blocks." in k or ".mid_block." in k for k in sd):
            sd = diffusers_vae_to_bfl(sd)
        own = {n for n, _ in self.named_parameters()}
        if not strict:
            sd = {k: v for k, v in sd.items() if k in own}
        else:
            sd = {k: v for k, v in sd.items() if k.startswith(prefix)}
        res = super().load_state_dict(sd, strict=strict, assign=assign)
        self._initialized |= own & set(sd.keys())
        return res

    def mark_initialized(self) -> None:
        self._initialized = {n for n, _ in self.named_parameters()}

    def uninitialized(self) -> list[str]:
        return [n for n, _ in self.named_parameters() if n not in self._initialized]

    def _require_weights(self) -> None:
        missing = self.uninitialized()
        if missing:
            raise _lib.VcbError(f"{len(missing)} VAE parameters were never loaded (e.g. {missing[:3]}): pass ae_ckpt= / call "
                                "load_state_dict or init_synthetic() first")

    def init_synthetic(self, seed: int = 0) -> "AutoEncoderDecoder":
        dev = next(self.parameters()).device
        g = torch.Generator(device=dev).manual_seed(seed)
        with torch.no_grad():
            for name, p in self.named_parameters():
                if p.dim() == 4:
                    p.copy_(torch.randn(p.shape, generator=g, device=dev) / math.sqrt(p.shape[1] * p.shape[2] * p.shape[3]))
                elif ".norm" in name and name.endswith(".weight"):
                    p.copy_(1.0 + 0.1 * torch.randn(p.shape, generator=g, device=dev))
                else:
                    p.copy_(0.02 * torch.randn(p.shape, generator=g, device=dev))
        self._key = None
        self.mark_initialized()
        return self

    # ---- packing --------------------------------------------------------------------------------
    def _conv(self, name: str, pad_out_to: int = 8) -> ConvW:
        w, b = self._p[name + ".weight"], self._p[name + ".bias"]
        co, ci, k, _ = w.shape
        if k == 3:
            cip = (ci + 63) // 64 * 64
            cop = (co + pad_out_to - 1) // pad_out_to * pad_out_to
            wp = torch.zeros(cop, 3, 3, cip, dtype=BF16, device=w.device)
            wp[:co, :, :, :ci] = w.permute(0, 2, 3, 1).to(BF16)           # [cout, ky, kx, cin]: tap-major, channels last
            wp = wp.reshape(cop, 9 * cip).contiguous()
            bp = torch.zeros(cop, dtype=torch.float32, device=w.device)
            bp[:co] = b.float()
            self._keep += [wp, bp]
            return ConvW(wp.data_ptr(), bp.data_ptr(), cip, cop)
        wp = w.reshape(co, ci).to(BF16).contiguous()
        bp = b.float().contiguous()
        self._keep += [wp, bp]
        return ConvW(wp.data_ptr(), bp.data_ptr(), ci, co)

    def _gn(self, name: str) -> GnW:
        g, b = self._p[name + ".weight"].float().contiguous(), self._p[name + ".bias"].float().contiguous()
        self._keep += [g, b]
        return GnW(g.data_ptr(), b.data_ptr())

    def _res(self, name: str) -> ResblockW:
        r = ResblockW()
        r.norm1, r.conv1 = self._gn(name + ".norm1"), self._conv(name + ".conv1")
        r.norm2, r.conv2 = self._gn(name + ".norm2"), self._conv(name + ".conv2")
        if name + ".nin_shortcut.weight" in self._p:
            r.shortcut = self._conv(name + ".nin_shortcut")
        return r

    def _engine(self):
        self._require_weights()
        key = tuple((p.data_ptr(), p._version) for p in self.parameters())
        if self._h is not None and key == self._key:
            return self._h
        some = next(self.parameters())
        if not some.is_cuda:
            raise _lib.VcbError("the VAE decoder runs on a CUDA device only (no CPU fallback)")
        lib = _lib.lib()
        if self._h is not None:
            lib.vcb_vae_destroy(self._h)
        P = self.params
        self._p = dict(self.named_parameters())
        self._keep = []
        with torch.no_grad():
            cfg = VaeConfigC()
            cfg.ch, cfg.out_ch, cfg.z_channels = P.ch, P.out_ch, P.z_channels
            cfg.num_res_blocks, cfg.n_levels = P.num_res_blocks, len(P.ch_mult)
            for i, m in enumerate(P.ch_mult):
                cfg.ch_mult[i] = m
            cfg.scale_factor, cfg.shift_factor = P.scale_factor, P.shift_factor
            w = VaeWeightsC()
            w.conv_in = self._conv("decoder.conv_in")
            w.mid1, w.mid2 = self._res("decoder.mid.block_1"), self._res("decoder.mid.block_2")
            w.attn_norm = self._gn("decoder.mid.attn_1.norm")
            w.attn_q, w.attn_k = self._conv("decoder.mid.attn_1.q"), self._conv("decoder.mid.attn_1.k")
            w.attn_v, w.attn_proj = self._conv("decoder.mid.attn_1.v"), self._conv("decoder.mid.attn_1.proj_out")
            nl = len(P.ch_mult)
            ups = (ResblockW * (nl * (P.num_res_blocks + 1)))()
            upc = (ConvW * max(1, nl - 1))()
            bi = ui = 0
            for lvl in reversed(range(nl)):
                for b in range(P.num_res_blocks + 1):
                    ups[bi] = self._res(f"decoder.up.{lvl}.block.{b}")
                    bi += 1
                if lvl != 0:
                    upc[ui] = self._conv(f"decoder.up.{lvl}.upsample.conv")
                    ui += 1
            w.up_blocks, w.upsample = C.cast(ups, C.POINTER(ResblockW)), C.cast(upc, C.POINTER(ConvW))
            w.norm_out, w.conv_out = self._gn("decoder.norm_out"), self._conv("decoder.conv_out")
        self._cw = (cfg, w, ups, upc)
        h = C.c_void_p()
        check(lib.vcb_vae_create(C.byref(cfg), C.byref(w), C.byref(h)), "vcb_vae_create")
        self._h, self._key = h, key
        return h

    # ---- decode -----------------------------------------------------------------------------------
    def _run(self, tokens: Tensor, h: int, w: int, want_raw: bool, want_img: bool):
        hnd = self._engine()
        lib = _lib.lib()
        n = tokens.shape[0]
        dev = tokens.device
        tokens = tokens.to(BF16).contiguous()
        up = 2 ** len(self.params.ch_mult)          # 2 (un-patchify) * 2^(levels-1)
        H, W = h * up, w * up
        need = lib.vcb_vae_workspace_bytes(hnd, n, h, w)
        if self._ws is None or self._ws.numel() < need or self._ws.device != dev:
            self._ws = torch.empty(int(need), dtype=torch.uint8, device=dev)
        raw = torch.empty(n, self.params.out_ch, H, W, dtype=torch.float32, device=dev) if want_raw else None
        img = torch.empty(n, self.params.out_ch, H, W, dtype=torch.uint8, device=dev) if want_img else None
        check(lib.vcb_vae_decode(hnd, self._ws.data_ptr(), self._ws.numel(), tokens.data_ptr(), n, h, w,
                                 None if raw is None else raw.data_ptr(), None if img is None else img.data_ptr(),
                                 torch.cuda.current_stream().cuda_stream), "vcb_vae_decode")
        return raw, img

    def decode_packed(self, tokens: Tensor, h: int, w: int) -> Tensor:
        """packed latent tokens [n, h*w, 4*z] (the sampler's output rows) -> uint8 image [n, 3, 16h, 16w]:
        un-patchify, ``z / scale + shift``, decoder, ``(x + 1) / 2``, clamp, ``to_pil_image`` scaling
        (visualcloze.py:428-439) in one call."""
        return self._run(tokens, h, w, False, True)[1]

    def decode(self, z: Tensor) -> Tensor:
        """``AutoEncoder.decode`` (autoencoder.py:307-309): z [n, 16, 2h, 2w] -> image tensor [n, 3, 16h, 16w] (z's dtype)."""
        n, c, H2, W2 = z.shape
        h, w = H2 // 2, W2 // 2
        tok = z.reshape(n, c, h, 2, w, 2).permute(0, 2, 4, 1, 3, 5).reshape(n, h * w, c * 4)
        return self._run(tok, h, w, True, False)[0].to(z.dtype)


# --------------------------------------------------------------------------------------------------
# encoder ("next" row (f)-1): models/modules/autoencoder.py:109-180, 262-275, 302-305
# --------------------------------------------------------------------------------------------------
def encoder_param_shapes(p: AutoEncoderParams) -> dict[str, tuple]:
    sh: dict[str, tuple] = {}

    def conv(n, ci, co, k):
        sh[n + ".weight"], sh[n + ".bias"] = (co, ci, k, k), (co,)

    def norm(n, c):
        sh[n + ".weight"], sh[n + ".bias"] = (c,), (c,)

    def res(n, ci, co):
        norm(n + ".norm1", ci); conv(n + ".conv1", ci, co, 3); norm(n + ".norm2", co); conv(n + ".conv2", co, co, 3)
        if ci != co:
            conv(n + ".nin_shortcut", ci, co, 1)

    conv("encoder.conv_in", p.in_channels, p.ch, 3)
    in_mult = (1,) + tuple(p.ch_mult)
    bi = p.ch
    for lvl in range(len(p.ch_mult)):
        bi, bo = p.ch * in_mult[lvl], p.ch * p.ch_mult[lvl]
        for b in range(p.num_res_blocks):
            res(f"encoder.down.{lvl}.block.{b}", bi, bo)
            bi = bo
        if lvl != len(p.ch_mult) - 1:
            conv(f"encoder.down.{lvl}.downsample.conv", bi, bi, 3)
    res("encoder.mid.block_1", bi, bi)
    norm("encoder.mid.attn_1.norm", bi)
    for n in ("q", "k", "v", "proj_out"):
        conv(f"encoder.mid.attn_1.{n}", bi, bi, 1)
    res("encoder.mid.block_2", bi, bi)
    norm("encoder.norm_out", bi)
    conv("encoder.conv_out", bi, 2 * p.z_channels, 3)
    return sh


class AutoEncoderEncoder(AutoEncoderDecoder):
    """VAE encoder with the reference's ``encoder.*`` keys.  ``encode_packed`` returns what the pipeline needs as
    ``fill_cond``: patchified ``(sample - shift) * scale`` tokens (visualcloze.py:377-388) in one call."""

    def __init__(self, params: AutoEncoderParams | None = None, device=None, dtype=BF16):
        nn.Module.__init__(self)
        self.params = params or AutoEncoderParams()
        self.scale_factor, self.shift_factor = self.params.scale_factor, self.params.shift_factor
        for name, shp in encoder_param_shapes(self.params).items():
            _attach(self, name, nn.Parameter(torch.empty(shp, device=device, dtype=dtype), requires_grad=False))
        self._h = None
        self._key = None
        self._ws = None
        self._initialized = set()

    def _engine(self):
        self._require_weights()
        key = tuple((p.data_ptr(), p._version) for p in self.parameters())
        if self._h is not None and key == self._key:
            return self._h
        some = next(self.parameters())
        if not some.is_cuda:
            raise _lib.VcbError("the VAE encoder runs on a CUDA device only (no CPU fallback)")
        lib = _lib.lib()
        if self._h is not None:
            lib.vcb_vae_enc_destroy(self._h)
        P = self.params
        self._p = dict(self.named_parameters())
        self._keep = []
        with torch.no_grad():
            cfg = VaeConfigC()
            cfg.ch, cfg.out_ch, cfg.z_channels = P.ch, P.out_ch, P.z_channels
            cfg.num_res_blocks, cfg.n_levels = P.num_res_blocks, len(P.ch_mult)
            for i, m in enumerate(P.ch_mult):
                cfg.ch_mult[i] = m
            cfg.scale_factor, cfg.shift_factor = P.scale_factor, P.shift_factor
            w = VaeEncWeightsC()
            w.conv_in = self._conv("encoder.conv_in")
            nl = len(P.ch_mult)
            dn = (ResblockW * (nl * P.num_res_blocks))()
            dc = (ConvW * max(1, nl - 1))()
            bi = 0
            for lvl in range(nl):
                for b in range(P.num_res_blocks):
                    dn[bi] = self._res(f"encoder.down.{lvl}.block.{b}")
                    bi += 1
                if lvl != nl - 1:
                    dc[lvl] = self._conv(f"encoder.down.{lvl}.downsample.conv")
            w.down_blocks, w.downsample = C.cast(dn, C.POINTER(ResblockW)), C.cast(dc, C.POINTER(ConvW))
            w.mid1, w.mid2 = self._res("encoder.mid.block_1"), self._res("encoder.mid.block_2")
            w.attn_norm = self._gn("encoder.mid.attn_1.norm")
            w.attn_q, w.attn_k = self._conv("encoder.mid.attn_1.q"), self._conv("encoder.mid.attn_1.k")
            w.attn_v, w.attn_proj = self._conv("encoder.mid.attn_1.v"), self._conv("encoder.mid.attn_1.proj_out")
            w.norm_out, w.conv_out = self._gn("encoder.norm_out"), self._conv("encoder.conv_out")
        self._cw = (cfg, w, dn, dc)
        h = C.c_void_p()
        check(lib.vcb_vae_enc_create(C.byref(cfg), C.byref(w), C.byref(h)), "vcb_vae_enc_create")
        self._h, self._key = h, key
        return h

    def encode_packed(self, image: Tensor, noise: Tensor | None = None, return_moments: bool = False):
        """image [n, 3, H, W] in [-1, 1] -> condition tokens [n, (H/16)(W/16), 64] bf16; ``noise`` [n, 16, H/8, W/8] is the
        standard-normal draw of ``latent_dist.sample()`` (None = distribution mode)."""
        hnd = self._engine()
        lib = _lib.lib()
        n, _, H, W = image.shape
        dev = image.device
        img = image.float().contiguous()
        f = 2 ** (len(self.params.ch_mult) - 1)
        need = lib.vcb_vae_enc_workspace_bytes(hnd, n, H, W)
        if need < 0:
            raise ValueError(f"image size must be a multiple of {2 * f}, got {(H, W)}")
        if self._ws is None or self._ws.numel() < need or self._ws.device != dev:
            self._ws = torch.empty(int(need), dtype=torch.uint8, device=dev)
        zc = self.params.z_channels
        tok = torch.empty(n, (H // (2 * f)) * (W // (2 * f)), 4 * zc, dtype=BF16, device=dev)
        mom = torch.empty(n, 2 * zc, H // f, W // f, dtype=torch.float32, device=dev) if return_moments else None
        nz = None if noise is None else noise.float().contiguous()
        check(lib.vcb_vae_encode(hnd, self._ws.data_ptr(), self._ws.numel(), img.data_ptr(), n, H, W,
                                 None if nz is None else nz.data_ptr(), tok.data_ptr(), None if mom is None else mom.data_ptr(),
                                 torch.cuda.current_stream().cuda_stream), "vcb_vae_encode")
        return (tok, mom) if return_moments else tok

    def decode(self, *a, **k):          # not a decoder
        raise AttributeError("AutoEncoderEncoder has no decode()")

    decode_packed = decode
