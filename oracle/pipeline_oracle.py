"""Oracle for the pipeline stages around the sampler -- TEST INFRASTRUCTURE ONLY.

Restates, from the already resized / cropped cells on (the PIL host logic of visualcloze.py:300-360 is covered separately by
tests/test_host_cpu.py), what ``VisualClozeModel.process_images`` (visualcloze.py:363-467) and ``.upsampling`` (:147-245) do:

  grid rows + fill masks (:365-375) -> VAE encode, (z - shift) * scale in the AE dtype (:377-378, :200-203) -> mask rearranges
  (:381-382, :207-208) -> img_cond (:384-389) -> seeded noise rows on the device generator (:392-399) -> prepare_modified
  (the UNMODIFIED reference function from oracle/_ref) -> Euler sampler (oracle/sampler_oracle.py over oracle/flux_oracle.py)
  -> per-row unpatchify + VAE decode + (x + 1) / 2 + clamp + to_pil_image (:424-439) -> crop of the masked cells (:449-465),
  SDEdit blend ``img * (1 - s) + latent * s`` (:221) for the second stage.

Runs on whatever device the inputs live on; the encoders (t5 / clip / VAE encode) are the caller's callables, exactly as in
the product pipeline under test, so both sides see identical bits.
"""
from __future__ import annotations

import numpy as np
import torch
from einops import rearrange
from PIL import Image

from oracle import flux_oracle as fo
from oracle import sampler_oracle as so
from oracle import vae_oracle as vo

BF16 = torch.bfloat16


def image_transform(img: Image.Image) -> torch.Tensor:
    """T.Compose([ToTensor(), Normalize(0.5, 0.5)]) (visualcloze.py:133-137)."""
    a = torch.from_numpy(np.asarray(img.convert("RGB"), dtype=np.uint8).copy()).permute(2, 0, 1).float() / 255.0
    return (a - 0.5) / 0.5


def to_pil_image(x: torch.Tensor) -> Image.Image:
    """torchvision.transforms.functional.to_pil_image for a float CHW tensor in [0, 1]: mul(255).byte()."""
    return Image.fromarray(x.mul(255).byte().permute(1, 2, 0).cpu().numpy())


def _ref_prepare_modified():
    from oracle import ref_runner as rr
    rr._import()
    import models.sampling as ref_sampling          # noqa: E402  (oracle/_ref, unmodified)
    return ref_sampling.prepare_modified


def _pack_latent(lat: torch.Tensor, shift: float, scale: float) -> torch.Tensor:
    lat = (lat - shift) * scale                      # in the AE dtype (bf16), python-float scalars
    return rearrange(lat.to(BF16), "b c (h ph) (w pw) -> b (h w) (c ph pw)", ph=2, pw=2)


def _pack_mask(mask: torch.Tensor) -> torch.Tensor:
    mask = rearrange(mask, "b c (h ph) (w pw) -> b (c ph pw) h w", ph=8, pw=8)
    return rearrange(mask, "b c (h ph) (w pw) -> b (h w) (c ph pw)", ph=2, pw=2)


def _decode_row(vae_p, vae_cfg, tokens, lh, lw):
    z = rearrange(tokens, "b (h w) (c ph pw) -> b c (h ph) (w pw)", ph=2, pw=2, h=lh // 2, w=lw // 2)
    x = vo.decode(vae_p, vae_cfg, z.float())         # vo.decode applies z / scale + shift itself (autoencoder.py:307-309)
    x = ((x + 1.0) / 2.0).clamp(0.0, 1.0)
    return to_pil_image(x[0].float())


def _model_fn(flux_p, flux_cfg):
    def fn(inp, timesteps, **k):
        dev = inp.device
        return fo.flux_forward(flux_p, flux_cfg, img=inp, timesteps=timesteps.to(dev), **k, mode="cuda_bf16")
    return fn


@torch.no_grad()
def process_images(cells, mask_position, grid_h, grid_w, prompts, seed, cfg, steps, *, t5, clip, encode, flux_p, flux_cfg, vae_p,
                   vae_cfg, device, time_shifting_factor=1, taps: dict | None = None):
    """cells: grid_h * grid_w processed PIL cells (row-major).  Returns (list of cropped PIL results, query-row PIL, rng)."""
    shift, scale = vae_cfg.shift_factor, vae_cfg.scale_factor
    grid_image, fill_mask = [], []
    for i in range(grid_h):
        row = [image_transform(im) for im in cells[i * grid_w:(i + 1) * grid_w]]
        marks = mask_position if i == grid_h - 1 else [0] * len(mask_position)
        fill_mask.append(torch.cat([torch.full((1, 1, row[0].shape[1], row[0].shape[2]), fill_value=float(m), device=device) for m in marks], dim=3))
        grid_image.append(torch.cat(row, dim=2).to(device))
    fill_cond = torch.cat([_pack_latent(encode(img[None].to(BF16)), shift, scale) for img in grid_image], dim=1)
    fill_mask = torch.cat([_pack_mask(m) for m in fill_mask], dim=1)
    img_cond = torch.cat((fill_cond, fill_mask.to(BF16)), dim=-1)
    rng = torch.Generator(device=device).manual_seed(int(seed))
    noise, sizes = [], []
    for sub in grid_image:
        h, w = sub.shape[-2:]
        sizes.append((h, w))
        noise.append(torch.randn([1, 16, h // 8, w // 8], device=device, generator=rng).to(BF16))
    inp = _ref_prepare_modified()(t5=t5, clip=clip, img=[noise], prompt=[" ".join(prompts)], proportion_empty_prompts=0.0)
    kw = dict(txt=inp["txt"], txt_ids=inp["txt_ids"], txt_mask=inp["txt_mask"], y=inp["vec"], img_ids=inp["img_ids"],
              img_mask=inp["img_mask"], cond=img_cond, guidance=torch.full((1,), cfg, device=device, dtype=BF16))
    if taps is not None:
        taps.update(x=inp["img"], **kw)
    samples = so.sample_ode(inp["img"], _model_fn(flux_p, flux_cfg), kw, num_steps=int(steps), do_shift=True,
                            time_shifting_factor=time_shifting_factor)[-1][:1]
    if taps is not None:
        taps["latent"] = samples
    outs, start = [], 0
    for (h, w) in sizes:
        end = start + h * w // 256
        outs.append(_decode_row(vae_p, vae_cfg, samples[:, start:end, :], h // 8, w // 8))
        start = end
    query = outs[-1]
    ret = [query.crop((j * query.width // grid_w, 0, (j + 1) * query.width // grid_w, query.height))
           for j in range(grid_w) if mask_position[j]]
    return ret, query, rng


@torch.no_grad()
def upsampling(image: Image.Image, target_size, cfg, upsampling_steps, upsampling_noise, generator, content_prompt, *, t5, clip,
               encode, flux_p, flux_cfg, vae_p, vae_cfg, device, taps: dict | None = None):
    """visualcloze.py:147-245 after the prompt clean-up (the content-instruction table is host string logic)."""
    if target_size is None:
        target_size = (1024, 1024)
    if target_size[0] * target_size[1] > 1024 * 1024:
        ar = target_size[0] / target_size[1]
        new_h = int((1024 * 1024 / ar) ** 0.5)
        target_size = (int(new_h * ar), new_h)
    image = image.resize(((target_size[0] // 16) * 16, (target_size[1] // 16) * 16))
    if upsampling_noise >= 1.0:
        return image
    shift, scale = vae_cfg.shift_factor, vae_cfg.scale_factor
    x = image_transform(image).to(device)
    blank = torch.zeros_like(x, dtype=BF16)
    mask = torch.full((1, 1, x.shape[1], x.shape[2]), fill_value=1, device=device, dtype=BF16)
    latent = _pack_latent(encode(x[None].to(BF16)), shift, scale)
    blank = _pack_latent(encode(blank[None]), shift, scale)
    lh, lw = x.shape[1] // 8, x.shape[2] // 8
    img_cond = torch.cat((blank, _pack_mask(mask)), dim=-1)
    noise = torch.randn([1, 16, lh, lw], device=device, generator=generator).to(BF16)
    inp = _ref_prepare_modified()(t5=t5, clip=clip, img=[[noise]], prompt=[content_prompt], proportion_empty_prompts=0.0)
    x_t = inp["img"] * (1 - upsampling_noise) + latent * upsampling_noise
    kw = dict(txt=inp["txt"], txt_ids=inp["txt_ids"], txt_mask=inp["txt_mask"], y=inp["vec"], img_ids=inp["img_ids"],
              img_mask=inp["img_mask"], cond=img_cond, guidance=torch.full((1,), cfg, device=device, dtype=BF16))
    if taps is not None:
        taps.update(x=x_t, **kw)
    sample = so.sample_ode(x_t, _model_fn(flux_p, flux_cfg), kw, num_steps=int(upsampling_steps), do_shift=False,
                           time_shifting_factor=1.0, strength=upsampling_noise)[-1][:1]
    return _decode_row(vae_p, vae_cfg, sample, lh, lw)
