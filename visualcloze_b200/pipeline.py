"""``VisualClozeModel`` -- the reference's inference entry point on the B200-native hot path.

Mirror of ``visualcloze.py:78-467``: same constructor arguments, ``set_grid_size``, ``process_images`` and ``upsampling``
signatures and return types (list of PIL images / PIL image), same image preparation, mask/cond packing, seeding,
row-wise RoPE ids, sampler settings and output cropping.  What differs is where the work runs:

  * the flow model is ``visualcloze_b200.model.FluxLoraWrapper`` (libvcb200 kernels), driven through
    ``visualcloze_b200.transport.Sampler`` -- one ``vcb_flux_forward`` + one ``vcb_euler_update`` per step;
  * only the query (last) row is decoded, fused with un-patchify, scale/shift, ``(x + 1) / 2``, clamp and uint8
    conversion (the reference decodes every row and returns crops of the last one, visualcloze.py:424-453);
  * text encoders (T5-XXL / CLIP-L) are outside this path: they are injected as callables --
    ``t5(list[str]) -> [B, 512, 4096]``, ``clip(list[str]) -> [B, 768]`` -- e.g. the reference's own ``load_t5`` /
    ``load_clip``;
  * the VAE *encode* of the condition rows (visualcloze.py:377-388) runs on ``vae.AutoEncoderEncoder`` (libvcb200; the
    "next" row (f)-1 of SURVEY.md section 8) unless an external ``encode(image [1, 3, H, W]) -> latent [1, 16, H/8, W/8]``
    callable (sampled, un-scaled -- e.g. ``AutoencoderKL.encode(...).latent_dist.sample()``) is injected.
"""
from __future__ import annotations

import random

import numpy as np
import torch
from PIL import Image

from .model import FluxLoraWrapper, flux_dev_fill_params
from .sampling import prepare_modified
from .transport import Sampler, create_transport
from .vae import AutoEncoderDecoder, AutoEncoderEncoder

_CONTENT_PREFIXES = (
    "The content of the last image in the final row is: ", "The last image of the last row depicts: ",
    "In the final row, the last image shows: ", "The last image in the bottom row illustrates: ",
    "The content of the bottom-right image is: ", "The final image in the last row portrays: ",
    "The last image of the final row displays: ", "In the last row, the final image captures: ",
    "The bottom-right corner image presents: ", "The content of the last image in the concluding row is: ",
    "In the last row, ", "The editing instruction in the last row is: ",
)


def resize_with_aspect_ratio(img: Image.Image, resolution: int, divisible: int = 16, aspect_ratio=None) -> Image.Image:
    """area ~= resolution^2, aspect kept, both sides multiples of 16 (visualcloze.py:28-75, PIL branch)."""
    w, h = img.size
    if aspect_ratio is None:
        aspect_ratio = w / h
    new_h = int((resolution * resolution / aspect_ratio) ** 0.5)
    new_w = int(new_h * aspect_ratio)
    new_w = max(new_w // divisible, 1) * divisible
    new_h = max(new_h // divisible, 1) * divisible
    return img.resize((new_w, new_h), Image.LANCZOS)


def center_crop(image: Image.Image, target_size) -> Image.Image:
    w, h = image.size
    nw, nh = target_size
    left, top = (w - nw) // 2, (h - nh) // 2
    return image.crop((left, top, left + nw, top + nh))


def to_rgb_if_rgba(img: Image.Image) -> Image.Image:
    """util/imgproc.py:90-96: composite RGBA over white."""
    if img.mode.upper() == "RGBA":
        rgb = Image.new("RGB", img.size, (255, 255, 255))
        rgb.paste(img, mask=img.split()[3])
        return rgb
    return img


def image_transform(img: Image.Image) -> torch.Tensor:
    """ToTensor + Normalize(0.5, 0.5): uint8 HWC -> float CHW in [-1, 1] (visualcloze.py:133-137)."""
    a = np.asarray(to_rgb_if_rgba(img).convert("RGB"), dtype=np.uint8)
    t = torch.from_numpy(a.copy()).permute(2, 0, 1).float().div(255.0)
    return (t - 0.5) / 0.5


def _pack_mask(mask: torch.Tensor) -> torch.Tensor:
    """[1,1,H,W] -> "b c (h 8)(w 8) -> b (c 8 8) h w" -> "b c (h 2)(w 2) -> b (h w) (c 2 2)" (visualcloze.py:381-382)."""
    b, c, H, W = mask.shape
    m = mask.reshape(b, c, H // 8, 8, W // 8, 8).permute(0, 1, 3, 5, 2, 4).reshape(b, c * 64, H // 8, W // 8)
    h, w = H // 16, W // 16
    return m.reshape(b, c * 64, h, 2, w, 2).permute(0, 2, 4, 1, 3, 5).reshape(b, h * w, c * 256)


def _patchify(lat: torch.Tensor) -> torch.Tensor:
    """"b c (h ph)(w pw) -> b (h w)(c ph pw)", ph = pw = 2."""
    b, c, H, W = lat.shape
    return lat.reshape(b, c, H // 2, 2, W // 2, 2).permute(0, 2, 4, 1, 3, 5).reshape(b, (H // 2) * (W // 2), c * 4)


def _load_file(path: str, device) -> dict:
    """safetensors (BFL / diffusers releases) or a torch checkpoint (the LoRA files written by train.py:690-705)."""
    if str(path).endswith(".safetensors"):
        from safetensors.torch import load_file
        return load_file(path, device=str(device))
    return torch.load(path, map_location=device)


def _expand_state_dict(model, sd: dict) -> dict:
    """``optionally_expand_state_dict`` (models/util.py:456-472): zero-expand tensors smaller than the model's parameter
    (FLUX.1-dev ``img_in`` 64 -> 384 input channels when a non-Fill base is given)."""
    own = dict(model.named_parameters())
    for name, t in list(sd.items()):
        p = own.get(name)
        if p is not None and tuple(t.shape) != tuple(p.shape):
            big = torch.zeros_like(p, device=t.device)
            big[tuple(slice(0, d) for d in t.shape)] = t
            sd[name] = big
    return sd


class VisualClozeModel:
    def __init__(self, model_path=None, model_name="flux-dev-fill-lora", max_length=512, lora_rank=256, atol=1e-6, rtol=1e-3,
                 solver="euler", time_shifting_factor=1, resolution=384, precision="bf16", *, model=None, ae_decoder=None,
                 ae_encoder=None, t5=None, clip=None, encode=None, device=None, flux_ckpt=None, ae_ckpt=None):
        """Positional arguments as the reference (visualcloze.py:79-82).  The reference constructor also pulls the base
        FLUX.1-Fill-dev weights (``load_flow_model``: ``$FLUX_DEV_FILL`` or a hub download) and the FLUX VAE
        (``AutoencoderKL.from_pretrained``); there is no network on the inference box, so here they come from files:
        ``flux_ckpt`` (default ``$FLUX_DEV_FILL``, BFL safetensors or a torch file) and ``ae_ckpt`` (default ``$AE``; BFL
        ``ae.safetensors`` or diffusers ``vae/diffusion_pytorch_model.safetensors`` -- keys are converted).  Pre-built
        ``model`` / ``ae_decoder`` / ``ae_encoder`` objects are used as they are.  Nothing is ever run on never-loaded
        parameters: the engines refuse to pack them."""
        if precision != "bf16":
            raise NotImplementedError("the sm_100a kernels implement the reference's default bf16 path only")
        if model_name != "flux-dev-fill-lora":
            raise NotImplementedError("VisualCloze uses the flux-dev-fill-lora geometry (models/util.py:132-165)")
        self.atol, self.rtol, self.solver = atol, rtol, solver
        self.time_shifting_factor, self.resolution, self.precision = time_shifting_factor, resolution, precision
        self.max_length, self.lora_rank = max_length, lora_rank
        self.device = torch.device(device if device is not None else "cuda")
        self.dtype = torch.bfloat16
        import os
        flux_ckpt = flux_ckpt or os.getenv("FLUX_DEV_FILL")
        ae_ckpt = ae_ckpt or os.getenv("AE")
        if model is None:
            with torch.device(self.device):
                model = FluxLoraWrapper(lora_rank=lora_rank, params=flux_dev_fill_params())
            if flux_ckpt is not None:                               # load_flow_model (models/util.py:406-411)
                sd = _expand_state_dict(model, _load_file(flux_ckpt, self.device))
                res = model.load_state_dict(sd, strict=False)
                if res.unexpected_keys:
                    raise ValueError(f"{flux_ckpt}: {len(res.unexpected_keys)} keys do not belong to the flux-dev-fill geometry, "
                                     f"e.g. {res.unexpected_keys[:3]}")
                del sd
        self.model = model
        if model_path is not None:
            ckpt = _load_file(model_path, self.device)
            res = self.model.load_state_dict(ckpt, strict=False)    # LoRA tensors on top of the base (visualcloze.py:111-112)
            if res.unexpected_keys:
                raise ValueError(f"{model_path}: unexpected keys, e.g. {res.unexpected_keys[:3]}")
            del ckpt
        ae_sd = _load_file(ae_ckpt, self.device) if ae_ckpt is not None and (ae_decoder is None or (ae_encoder is None and encode is None)) else None
        if ae_decoder is None:
            with torch.device(self.device):
                ae_decoder = AutoEncoderDecoder()
            if ae_sd is not None:
                ae_decoder.load_state_dict(ae_sd, strict=True)
        self.ae = ae_decoder
        if ae_encoder is None and encode is None:
            with torch.device(self.device):
                ae_encoder = AutoEncoderEncoder()
            if ae_sd is not None:
                ae_encoder.load_state_dict(ae_sd, strict=True)
        self.ae_encoder = ae_encoder
        # fail at construction, not at the first image: a drop-in constructor must not leave torch.empty weights behind
        for what, mod in (("FLUX.1-Fill base weights (flux_ckpt= / $FLUX_DEV_FILL)", self.model), ("VAE decoder (ae_ckpt= / $AE)", self.ae),
                          ("VAE encoder (ae_ckpt= / $AE, or pass encode=)", self.ae_encoder)):
            missing = mod.uninitialized() if (mod is not None and hasattr(mod, "uninitialized")) else []
            if missing:
                raise RuntimeError(f"VisualClozeModel: {what}: {len(missing)} parameters were never loaded (e.g. {missing[:3]})")
        self.t5, self.clip, self.encode = t5, clip, encode
        self.sampler = Sampler(create_transport("Linear", "velocity", do_shift=True))
        self.sample_fn = self._make_sample_fn(30, True, self.time_shifting_factor, None)
        self.image_transform = image_transform
        self.grid_h = None
        self.grid_w = None

    def _make_sample_fn(self, num_steps, do_shift, tsf, strength):
        return self.sampler.sample_ode(sampling_method=self.solver, num_steps=num_steps, atol=self.atol, rtol=self.rtol,
                                       reverse=False, do_shift=do_shift, time_shifting_factor=tsf, strength=strength)

    def set_grid_size(self, h, w):
        self.grid_h, self.grid_w = h, w

    def _need(self, what):
        fn = getattr(self, what)
        if fn is None:
            raise RuntimeError(f"VisualClozeModel needs a `{what}` callable (text encoders and the VAE encoder are outside the "
                               "hot path; pass the reference's own, see the module docstring)")
        return fn

    def _encode_tokens(self, image_chw: torch.Tensor) -> torch.Tensor:
        """patchify((ae.encode(x).latent_dist.sample() - shift) * scale) in bf16 (visualcloze.py:377-378,384-385): [1, hw/256, 64].
        Like the reference, the sampling noise comes from the global (unseeded) RNG."""
        if self.encode is not None:
            lat = self.encode(image_chw[None].to(self.device, self.dtype))
            return _patchify(((lat - self.ae.shift_factor) * self.ae.scale_factor).to(self.dtype))
        x = image_chw[None].to(self.device)
        noise = torch.randn(1, self.ae_encoder.params.z_channels, x.shape[2] // 8, x.shape[3] // 8, device=self.device)
        return self.ae_encoder.encode_packed(x, noise)

    # ------------------------------------------------------------------------------------------------
    def _prepare_grid(self, images):
        """Resize / crop every cell to its row's reference size; blank + mask for the missing targets of the last row
        (visualcloze.py:300-360)."""
        grid_h, grid_w, resolution = self.grid_h, self.grid_w, self.resolution
        processed, mask_position = [], []
        target_size = upsampling_size = None
        for i in range(grid_h):
            reference_size = None
            for j in range(grid_w):
                if images[i][j] is not None:
                    if i == grid_h - 1 and upsampling_size is None:
                        upsampling_size = images[i][j].size
                    reference_size = resize_with_aspect_ratio(images[i][j], resolution).size
                    if i == grid_h - 1 and target_size is None:
                        target_size = reference_size
                    break
            for j in range(grid_w):
                if images[i][j] is not None:
                    t = resize_with_aspect_ratio(images[i][j], resolution)
                    if t.width <= t.height:
                        t = t.resize((reference_size[0], int(reference_size[0] / t.width * t.height)))
                    else:
                        t = t.resize((int(reference_size[1] / t.height * t.width), reference_size[1]))
                    processed.append(center_crop(t, reference_size))
                    if i == grid_h - 1:
                        mask_position.append(0)
                else:
                    if i != grid_h - 1:
                        raise ValueError("Please provide each image in the in-context example.")
                    processed.append(Image.new("RGB", reference_size or (resolution, resolution), (0, 0, 0)))
                    mask_position.append(1)
        if len(mask_position) > 1 and sum(mask_position) > 1:
            new_w = 384 if target_size is None else target_size[0]
            for k, im in enumerate(processed):
                new_h = int(im.height * (new_w / im.width))
                new_w, new_h = int(new_w / 16) * 16, int(new_h / 16) * 16
                processed[k] = im.resize((new_w, new_h))
        return processed, mask_position, upsampling_size

    @torch.no_grad()
    def process_images(self, images, prompts, seed: int = 0, cfg: int = 30, steps: int = 30, upsampling_steps: int = 10,
                       upsampling_noise: float = 0.4, is_upsampling: bool = True):
        if seed == 0:
            seed = random.randint(0, 2 ** 32 - 1)
        self.sample_fn = self._make_sample_fn(int(steps), True, self.time_shifting_factor, None)
        grid_h, grid_w = self.grid_h, self.grid_w
        for i in range(grid_h):
            images[i] = [im.convert("RGB") if im is not None else None for im in images[i]]
        processed, mask_position, upsampling_size = self._prepare_grid(images)

        dev, dt = self.device, self.dtype
        rows, fill_mask = [], []
        for i in range(grid_h):
            cells = [self.image_transform(im) for im in processed[i * grid_w:(i + 1) * grid_w]]
            h, w = cells[0].shape[1], cells[0].shape[2]
            marks = mask_position if i == grid_h - 1 else [0] * grid_w
            rows.append(torch.cat(cells, dim=2).to(dev))
            fill_mask.append(torch.cat([torch.full((1, 1, h, w), float(m), device=dev) for m in marks], dim=3))
        fill_cond = torch.cat([self._encode_tokens(r) for r in rows], dim=1)
        fill_mask = torch.cat([_pack_mask(m) for m in fill_mask], dim=1).to(dt)
        img_cond = torch.cat((fill_cond, fill_mask), dim=-1)

        rng = torch.Generator(device=dev).manual_seed(int(seed))
        noise, sizes = [], []
        for r in rows:
            h, w = r.shape[-2:]
            sizes.append((h, w))
            noise.append(torch.randn([1, 16, h // 8, w // 8], device=dev, generator=rng).to(dt))
        inp = prepare_modified(t5=self._need("t5"), clip=self._need("clip"), img=[noise], prompt=[" ".join(prompts)],
                               proportion_empty_prompts=0.0)
        kw = dict(txt=inp["txt"], txt_ids=inp["txt_ids"], txt_mask=inp["txt_mask"], y=inp["vec"], img_ids=inp["img_ids"],
                  img_mask=inp["img_mask"], cond=img_cond, guidance=torch.full((1,), cfg, device=dev, dtype=dt))
        samples = self.sample_fn(inp["img"], self.model.forward, kw)[-1][:1]

        # decode the query row only (the reference decodes all rows and crops the last, visualcloze.py:420-453)
        qh, qw = sizes[-1]
        n_tok = qh * qw // 256
        start = sum(h * w // 256 for h, w in sizes[:-1])
        tile = self.ae.decode_packed(samples[:, start:start + n_tok, :], qh // 16, qw // 16)[0]     # uint8 [3, H, W]
        query = Image.fromarray(tile.permute(1, 2, 0).cpu().numpy())
        torch.cuda.empty_cache()

        ret = []
        ret_w, ret_h = query.width, query.height
        for j in range(grid_w):
            if not mask_position[j]:
                continue
            cropped = query.crop((j * ret_w // grid_w, 0, (j + 1) * ret_w // grid_w, ret_h))
            if is_upsampling:
                cropped = self.upsampling(cropped, upsampling_size, cfg, upsampling_steps=upsampling_steps,
                                          upsampling_noise=upsampling_noise, generator=rng, content_prompt=prompts[2])
            ret.append(cropped)
        return ret

    @torch.no_grad()
    def upsampling(self, image, target_size, cfg, upsampling_steps, upsampling_noise, generator, content_prompt):
        """SDEdit second stage (visualcloze.py:147-245): noise the resized image to `upsampling_noise`, denoise with an
        all-ones fill mask over a blank condition, 1x1 grid."""
        for c in _CONTENT_PREFIXES:
            if content_prompt.startswith(c):
                content_prompt = content_prompt.replace(c, "")
        if target_size is None:
            target_size = (1024, 1024)
        if target_size[0] * target_size[1] > 1024 * 1024:
            ar = target_size[0] / target_size[1]
            new_h = int((1024 * 1024 / ar) ** 0.5)
            target_size = (int(new_h * ar), new_h)
        image = image.resize(((target_size[0] // 16) * 16, (target_size[1] // 16) * 16))
        if upsampling_noise >= 1.0:
            return image
        self.sample_fn = self._make_sample_fn(int(upsampling_steps), False, 1.0, upsampling_noise)
        dev, dt = self.device, self.dtype
        x = self.image_transform(image).to(dev)
        latent = self._encode_tokens(x)
        blank = self._encode_tokens(torch.zeros_like(x))
        lh, lw = x.shape[1] // 8, x.shape[2] // 8
        mask = _pack_mask(torch.ones(1, 1, x.shape[1], x.shape[2], device=dev, dtype=dt))
        img_cond = torch.cat((blank, mask), dim=-1)
        noise = torch.randn([1, 16, lh, lw], device=dev, generator=generator).to(dt)
        inp = prepare_modified(t5=self._need("t5"), clip=self._need("clip"), img=[[noise]], prompt=[content_prompt],
                               proportion_empty_prompts=0.0)
        x_t = (inp["img"] * (1 - upsampling_noise) + latent * upsampling_noise).to(dt)
        kw = dict(txt=inp["txt"], txt_ids=inp["txt_ids"], txt_mask=inp["txt_mask"], y=inp["vec"], img_ids=inp["img_ids"],
                  img_mask=inp["img_mask"], cond=img_cond, guidance=torch.full((1,), cfg, device=dev, dtype=dt))
        sample = self.sample_fn(x_t, self.model.forward, kw)[-1][:1]
        tile = self.ae.decode_packed(sample, lh // 2, lw // 2)[0]
        return Image.fromarray(tile.permute(1, 2, 0).cpu().numpy())
