"""Pins oracle/ against outputs of the reference's own modules (tests/golden/, made by
oracle/gen_golden.py in the build container).  CPU only."""
import os

import pytest
import torch

from oracle import flux_oracle as fo
from oracle import sampler_oracle as so
from oracle import vae_oracle as vo
from conftest import GOLDEN, rel_l2


def _load(name):
    return torch.load(os.path.join(GOLDEN, name), weights_only=False)


@pytest.mark.parametrize("tag", ["b1", "b2r"])
def test_flux_forward_fp32_matches_reference(tag):
    g = _load(f"flux_small_{tag}.pt")
    cfg = fo.FluxConfig(**g["cfg"])
    p = fo.make_params(cfg, seed=g["param_seed"])
    out = fo.flux_forward(p, cfg, **g["inputs"], mode="fp32")
    # fp32 vs fp32: only summation-order noise
    assert rel_l2(out, g["out_fp32"]) < 2e-5
    assert out.shape == g["out_fp32"].shape


@pytest.mark.parametrize("tag", ["b1", "b2r"])
def test_flux_forward_cpu_bf16_matches_reference(tag):
    g = _load(f"flux_small_{tag}.pt")
    cfg = fo.FluxConfig(**g["cfg"])
    p = fo.make_params(cfg, seed=g["param_seed"])
    out = fo.flux_forward(p, cfg, **g["inputs"], mode="cpu_bf16")
    assert out.dtype == torch.bfloat16
    # same rounding points; SDPA-vs-explicit softmax differ by bf16 noise
    assert rel_l2(out, g["out_cpu_bf16"]) < 1.5e-2
    # the authoritative CUDA-autocast mode differs only by LayerNorm returning fp32: bf16 noise.
    # (fp32 is not comparable: bf16 guidance rounds 1000*30 to 29952 before the sinusoid.)
    out_c = fo.flux_forward(p, cfg, **g["inputs"], mode="cuda_bf16")
    assert rel_l2(out_c, g["out_cpu_bf16"]) < 2e-2


def test_ragged_mask_rows_are_zeroed_then_projected():
    g = _load("flux_small_b2r.pt")
    cfg = fo.FluxConfig(**g["cfg"])
    p = fo.make_params(cfg, seed=g["param_seed"])
    inp = dict(g["inputs"])
    out = fo.flux_forward(p, cfg, **inp, mode="fp32")
    # changing the content of padded img tokens of sample 1 must not change its valid rows
    img2 = inp["img"].clone()
    pad = inp["img_mask"][1] == 0
    img2[1, pad] = 7.0
    out2 = fo.flux_forward(p, cfg, **dict(inp, img=img2), mode="fp32")
    valid = ~pad
    assert torch.allclose(out[1, valid], out2[1, valid], atol=1e-5)
    assert torch.equal(out[0], out2[0])


@pytest.mark.parametrize("tag", ["shift4", "sdedit5"])
def test_sampler_matches_reference_transport(tag):
    g = _load("sampler.pt")[tag]
    cfg_g = _load("flux_small_b1.pt")
    cfg = fo.FluxConfig(**cfg_g["cfg"])
    p = fo.make_params(cfg, seed=cfg_g["param_seed"])
    seen = []

    def model_fn(x, timesteps, **kw):
        seen.append(timesteps.clone())
        return fo.flux_forward(p, cfg, img=x, timesteps=timesteps, **kw, mode="cpu_bf16")

    kw = dict(g["kwargs"])
    mk = dict(g["inputs"], cond=g["cond"])
    traj = so.sample_ode(g["x"], model_fn, mk, num_steps=kw["num_steps"], do_shift=kw["do_shift"],
                         time_shifting_factor=kw["time_shifting_factor"], strength=kw.get("strength"))
    assert "cond" in mk
    assert traj.shape == g["traj"].shape and traj.dtype == torch.bfloat16
    # FLUX-time fed to the model at every evaluation: N points -> N-1 evaluations
    assert torch.allclose(torch.stack(seen), g["flux_t"], atol=1e-6)
    assert len(seen) == kw["num_steps"] - 1
    assert torch.equal(traj[0], g["traj"][0])
    assert rel_l2(traj[-1], g["traj"][-1]) < 2e-2


def test_solver_grid_equals_get_schedule():
    g = _load("sampler.pt")["get_schedule"]
    for (n, L), ref in ((k, v) for k, v in g.items() if isinstance(k, tuple)):
        tau = so.solver_grid(n + 1, L, do_shift=True, time_shifting_factor=1)
        assert torch.allclose(1 - tau, torch.tensor(ref), atol=3e-7), (n, L)


def test_vae_decode_matches_reference():
    g = _load("vae_small.pt")
    cfg = vo.VaeConfig(**g["cfg"])
    p = vo.make_decoder_params(cfg, seed=g["param_seed"], dtype=torch.float32)
    out = vo.decode(p, cfg, g["z"])
    assert out.shape == g["out_fp32"].shape
    assert rel_l2(out, g["out_fp32"]) < 2e-5


def test_vae_encode_moments_match_reference():
    g = _load("vae_small.pt")
    cfg = vo.VaeConfig(**g["cfg"])
    p = vo.make_encoder_params(cfg, seed=g["enc_param_seed"], dtype=torch.float32)
    mom = vo.encode_moments(p, cfg, g["img"])
    assert mom.shape == g["moments_fp32"].shape
    assert rel_l2(mom, g["moments_fp32"]) < 2e-5
