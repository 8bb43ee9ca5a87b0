"""``VisualClozePipeline``-shaped front end (the diffusers pipeline the reference's README points to, README.md:141-205) over
``VisualClozeModel`` -- the caller on the far side of the hot path.

The diffusers pipeline is not part of /root/reference (README.md:143 links it); what is mirrored here is its documented call:

    pipe(task_prompt=..., content_prompt=..., image=[[...], [..., None]], upsampling_height=..., upsampling_width=...,
         upsampling_strength=..., guidance_scale=30, num_inference_steps=30, max_sequence_length=512, generator=...).images[0][0]

mapped onto the reference's own entry (visualcloze.py:247-467): ``image`` is the grid (rows of PIL images, ``None`` = target),
the layout prompt is the first template of ``get_layout_instruction`` (data/prefix_instruction.py:684-697, what app.py:116
pre-fills: "A grid layout with R rows and C columns, displaying N images arranged side by side."), ``upsampling_strength`` is the SDEdit
noise level, ``generator`` supplies the seed.  Text encoders / VAE weights are whatever the wrapped ``VisualClozeModel`` holds.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import torch


@dataclass
class PipelineOutput:
    images: list = field(default_factory=list)          # [sample][target] PIL images, like the diffusers pipeline output


def layout_prompt(rows: int, cols: int) -> str:
    """first template of the reference's ``get_layout_instruction(cols, rows)`` (data/prefix_instruction.py:686)"""
    return f"A grid layout with {rows} rows and {cols} columns, displaying {rows * cols} images arranged side by side."


class VisualClozePipelineAdapter:
    def __init__(self, model):
        self.model = model                                # a visualcloze_b200.pipeline.VisualClozeModel

    @torch.no_grad()
    def __call__(self, task_prompt: str, content_prompt: str | None, image: list, upsampling_height: int | None = None,
                 upsampling_width: int | None = None, upsampling_strength: float = 0.4, guidance_scale: float = 30.0,
                 num_inference_steps: int = 30, upsampling_steps: int = 10, max_sequence_length: int = 512, generator=None,
                 **unused) -> PipelineOutput:
        if not isinstance(image, list) or not image or not all(isinstance(r, list) for r in image):
            raise ValueError("image must be a list of rows (lists of PIL images; None marks a target in the last row)")
        cols = len(image[0])
        if any(len(r) != cols for r in image):
            raise ValueError("every grid row needs the same number of images")
        if any(im is None for r in image[:-1] for im in r):
            raise ValueError("Please provide each image in the in-context example.")
        if max_sequence_length != self.model.max_length:
            raise ValueError(f"the wrapped model was built for max_length={self.model.max_length}")
        seed = int(generator.initial_seed()) if generator is not None else 0
        self.model.set_grid_size(len(image), cols)
        prompts = [layout_prompt(len(image), cols), task_prompt or "", content_prompt or ""]
        is_up = upsampling_strength < 1.0 and (upsampling_height is not None or upsampling_width is not None)
        grid = [list(r) for r in image]
        out = self.model.process_images(grid, prompts, seed=seed, cfg=guidance_scale, steps=int(num_inference_steps),
                                        upsampling_steps=int(upsampling_steps), upsampling_noise=float(upsampling_strength),
                                        is_upsampling=is_up)
        if is_up and upsampling_height and upsampling_width:
            out = [im.resize((int(upsampling_width) // 16 * 16, int(upsampling_height) // 16 * 16)) if im.size != (upsampling_width, upsampling_height) else im
                   for im in out]
        return PipelineOutput(images=[out])
