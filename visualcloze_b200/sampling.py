"""Schedule API and token packing with the reference's names and signatures (``models/sampling.py``).

Host-side logic only (shapes, ids, schedules); tensors keep the device/dtype they arrive with.
  get_noise :18-35 · prepare_modified :37-118 · time_shift :300-301 · get_lin_function :304-309 ·
  get_schedule :312-328 · denoise :331-361 · unpack :364-372
"""
from __future__ import annotations

import math
import random
from typing import Callable

import torch
import torch.nn.functional as F
from torch import Tensor


def get_noise(num_samples: int, height: int, width: int, device: torch.device, dtype: torch.dtype, seed: int):
    return torch.randn(num_samples, 16, 2 * math.ceil(height / 16), 2 * math.ceil(width / 16), device=device,
                       dtype=dtype, generator=torch.Generator(device=device).manual_seed(seed))


def _patchify(x: Tensor) -> Tensor:
    """"c (h ph) (w pw) -> (h w) (c ph pw)" with ph = pw = 2."""
    c, h, w = x.shape
    return x.reshape(c, h // 2, 2, w // 2, 2).permute(1, 3, 0, 2, 4).reshape((h // 2) * (w // 2), c * 4)


def prepare_modified(t5, clip, img: list[list[Tensor]], prompt: str | list[str], proportion_empty_prompts: float = 0.1,
                     is_train: bool = True, text_emb: list[dict[str, Tensor]] = None) -> dict[str, Tensor]:
    """Pack a batch of grids (one list of row latents per sample) into right-padded token sequences."""
    assert isinstance(img, list) and all(isinstance(i, list) for i in img)
    bs = len(img)
    max_len = max(sum(i.shape[-2] * i.shape[-1] for i in rows) for rows in img) // 4
    dev = img[0][0].device
    img_mask = torch.zeros(bs, max_len, device=dev, dtype=torch.int32)
    padded_img, padded_ids = [], []
    for i, rows in enumerate(img):
        toks, ids = [], []
        for j, row in enumerate(rows):
            row = row.squeeze(0)
            c, h, w = row.shape
            rid = torch.zeros(h // 2, w // 2, 3)
            rid[..., 0] = j + 1                                   # every grid row gets its own RoPE index
            rid[..., 1] = rid[..., 1] + torch.arange(h // 2)[:, None]
            rid[..., 2] = rid[..., 2] + torch.arange(w // 2)[None, :]
            ids.append(rid.reshape(-1, 3))
            toks.append(_patchify(row))
        toks, ids = torch.cat(toks, dim=0), torch.cat(ids, dim=0)
        n = toks.shape[0]
        padded_img.append(F.pad(toks, (0, 0, 0, max_len - n)))
        padded_ids.append(F.pad(ids, (0, 0, 0, max_len - n)))
        img_mask[i, :n] = 1
    img_t = torch.stack(padded_img, dim=0)
    img_ids = torch.stack(padded_ids, dim=0)

    if isinstance(prompt, str):
        prompt = [prompt]
    bs = len(prompt)
    drop_mask = []
    for idx in range(bs):
        if random.random() < proportion_empty_prompts:
            prompt[idx] = ""
        elif isinstance(prompt[idx], list):
            prompt[idx] = random.choice(prompt[idx]) if is_train else prompt[idx][0]
        drop_mask.append(0 if prompt[idx] == "" else 1)
    drop_mask = torch.tensor(drop_mask, device=img_mask.device, dtype=img_mask.dtype)

    txt = torch.stack([e["txt"] for e in text_emb], dim=0).to(img_t.device) if t5 is None else t5(prompt)
    if txt.shape[0] == 1 and bs > 1:
        txt = txt.expand(bs, *txt.shape[1:])
    txt_ids = torch.zeros(bs, txt.shape[1], 3)
    txt_mask = torch.ones(bs, txt.shape[1], device=txt.device, dtype=torch.int32)
    vec = torch.stack([e["vec"] for e in text_emb], dim=0).to(img_t.device) if clip is None else clip(prompt)
    if vec.shape[0] == 1 and bs > 1:
        vec = vec.expand(bs, *vec.shape[1:])
    d = img_t.device
    return {"img": img_t, "img_ids": img_ids.to(d), "txt": txt.to(d), "txt_ids": txt_ids.to(d), "vec": vec.to(d),
            "img_mask": img_mask.to(d), "txt_mask": txt_mask.to(txt.device), "drop_mask": drop_mask.to(d)}


def time_shift(mu: float, sigma: float, t: Tensor):
    return math.exp(mu) / (math.exp(mu) + (1 / t - 1) ** sigma)


def get_lin_function(x1: float = 256, y1: float = 0.5, x2: float = 4096, y2: float = 1.15) -> Callable[[float], float]:
    m = (y2 - y1) / (x2 - x1)
    b = y1 - m * x1
    return lambda x: m * x + b


def get_schedule(num_steps: int, image_seq_len: int, base_shift: float = 0.5, max_shift: float = 1.15,
                 shift: bool = True) -> list[float]:
    timesteps = torch.linspace(1, 0, num_steps + 1)          # extra step for zero
    if shift:
        mu = get_lin_function(y1=base_shift, y2=max_shift)(image_seq_len)
        timesteps = time_shift(mu, 1.0, timesteps)
    return timesteps.tolist()


def denoise(model, img: Tensor, img_ids: Tensor, txt: Tensor, txt_ids: Tensor, vec: Tensor, timesteps: list[float],
            guidance: float = 4.0, img_cond: Tensor | None = None, txt_mask: Tensor | None = None,
            img_mask: Tensor | None = None):
    """BFL-style Euler loop over ``get_schedule`` (reference :331-361).  The reference version passes no masks and
    therefore cannot run against its own ``Flux``; here all-ones masks are the default."""
    from . import ops
    B, Li, _ = img.shape
    if txt_mask is None:
        txt_mask = torch.ones(B, txt.shape[1], dtype=torch.int32, device=img.device)
    if img_mask is None:
        img_mask = torch.ones(B, Li, dtype=torch.int32, device=img.device)
    guidance_vec = torch.full((B,), guidance, device=img.device, dtype=img.dtype)
    for t_curr, t_prev in zip(timesteps[:-1], timesteps[1:]):
        t_vec = torch.full((B,), t_curr, dtype=img.dtype, device=img.device)
        pred = model(img=torch.cat((img, img_cond), dim=-1) if img_cond is not None else img, img_ids=img_ids, txt=txt,
                     txt_ids=txt_ids, y=vec, timesteps=t_vec, txt_mask=txt_mask, img_mask=img_mask, guidance=guidance_vec)
        # img + (t_prev - t_curr) * pred (:358): a python-float scalar is an fp32 opmath scalar (NOT rounded to bf16 first, unlike
        # the 0-dim tensor dt of the torchdiffeq path): bf16(img + bf16(fp32(dt) * pred)).  The kernel computes
        # x + bf16(dt * (-v)); the sign moves into dt, which is exact.
        new = torch.empty_like(img)
        ops.euler_update(img.reshape(B * Li, -1), pred.reshape(B * Li, -1), -(t_prev - t_curr), new.reshape(B * Li, -1))
        img = new
    return img


def unpack(x: Tensor, height: int, width: int) -> Tensor:
    """"b (h w) (c ph pw) -> b c (h ph) (w pw)", h = ceil(height/16), w = ceil(width/16), ph = pw = 2."""
    h, w = math.ceil(height / 16), math.ceil(width / 16)
    b, _, d = x.shape
    c = d // 4
    return x.reshape(b, h, w, c, 2, 2).permute(0, 3, 1, 4, 2, 5).reshape(b, c, h * 2, w * 2)
