"""Flow-matching sampler with the reference's ``transport`` API for the path VisualCloze uses.

  create_transport  transport/__init__.py:4-62       Sampler.sample_ode  transport/transport.py:361-410
  ode grid          transport/integrators.py:79-120  time_shift          transport/utils.py:33-43
  velocity_ode      transport/transport.py:193-198   Euler               torchdiffeq.odeint(method="euler")

``Sampler(transport).sample_ode(...)`` returns ``fn(x, model, model_kwargs)`` which returns the whole trajectory
``[num_steps, B, L, C]`` (callers index ``[-1]``).  ``num_steps`` time points mean ``num_steps - 1`` model
evaluations.  When ``model`` is the bound ``forward`` of a ``visualcloze_b200.model.Flux`` the step-invariant work
of all evaluations is hoisted (engine.prepare) and each step is one ``vcb_flux_forward`` + one ``vcb_euler_update``;
any other callable is driven step by step through the same update kernel.

Only what the inference pipeline uses is implemented: Linear path, velocity prediction, fixed-grid Euler.
Training losses, SDE samplers, likelihood ODE, VP/GVP paths and adaptive solvers are out of scope.
"""
from __future__ import annotations

import enum
import math

import torch as th

from . import ops


class ModelType(enum.Enum):
    NOISE = enum.auto()
    SCORE = enum.auto()
    VELOCITY = enum.auto()


class PathType(enum.Enum):
    LINEAR = enum.auto()
    GVP = enum.auto()
    VP = enum.auto()


class WeightType(enum.Enum):
    NONE = enum.auto()
    VELOCITY = enum.auto()
    LIKELIHOOD = enum.auto()


def time_shift(mu: float, sigma: float, t: th.Tensor):
    """transport/utils.py:33-38: the mirrored (t=0 noise, t=1 data) form of the FLUX shift."""
    t = 1 - t
    t = math.exp(mu) / (math.exp(mu) + (1 / t - 1) ** sigma)
    t = 1 - t
    return t


def get_lin_function(x1: float = 256, y1: float = 0.5, x2: float = 4096, y2: float = 1.15):
    m = (y2 - y1) / (x2 - x1)
    b = y1 - m * x1
    return lambda x: m * x + b


class Transport:
    def __init__(self, *, model_type, path_type, loss_type="mse", train_eps=0, sample_eps=0, snr_type="uniform",
                 do_shift=True):
        if path_type != PathType.LINEAR or model_type != ModelType.VELOCITY:
            raise NotImplementedError("only the Linear path with velocity prediction is on the VisualCloze inference path")
        self.model_type, self.path_type = model_type, path_type
        self.loss_type, self.train_eps, self.sample_eps = loss_type, train_eps, sample_eps
        self.snr_type, self.do_shift = snr_type, do_shift

    def check_interval(self, train_eps, sample_eps, *, diffusion_form="SBDM", sde=False, reverse=False, eval=False,
                       last_step_size=0.0):
        t0, t1 = 0, 1          # Linear + velocity "is stable everywhere" (transport/transport.py:70-96)
        if reverse:
            t0, t1 = 1 - t0, 1 - t1
        return t0, t1


def create_transport(path_type="Linear", prediction="velocity", loss_weight=None, train_eps=None, sample_eps=None,
                     snr_type="uniform", loss_type="mse", do_shift=True):
    model_type = {"noise": ModelType.NOISE, "score": ModelType.SCORE}.get(prediction, ModelType.VELOCITY)
    ptype = {"Linear": PathType.LINEAR, "GVP": PathType.GVP, "VP": PathType.VP}[path_type]
    return Transport(model_type=model_type, path_type=ptype, loss_type=loss_type, train_eps=0, sample_eps=0,
                     snr_type=snr_type, do_shift=do_shift)


def solver_grid(t0, t1, num_steps: int, seq_len: int, do_shift: bool, time_shifting_factor) -> th.Tensor:
    """integrators.py:99-101 and :114-116 -- fp32 grid of solver time (0 = noise, 1 = data)."""
    t = th.linspace(t0, t1, int(num_steps))
    if time_shifting_factor:
        t = t / (t + time_shifting_factor - time_shifting_factor * t)
    if do_shift:
        mu = get_lin_function(y1=0.5, y2=1.15)(seq_len)
        t = time_shift(mu, 1.0, t)
    return t


def _is_native_forward(model) -> bool:
    from .model import Flux
    return isinstance(getattr(model, "__self__", None), Flux) and getattr(model, "__name__", "") == "forward"


class Sampler:
    """Sampler class for the transport model (ODE / Euler only)."""

    def __init__(self, transport: Transport):
        self.transport = transport

    def sample_ode(self, *, sampling_method="dopri5", num_steps=50, atol=1e-6, rtol=1e-3, reverse=False, do_shift=True,
                   time_shifting_factor=None, strength=None):
        if sampling_method != "euler":
            raise NotImplementedError("only the fixed-grid 'euler' solver (the pipeline default, visualcloze.py:81) is implemented")
        t0, t1 = self.transport.check_interval(self.transport.train_eps, self.transport.sample_eps, sde=False, eval=True,
                                               reverse=reverse, last_step_size=0.0)
        if strength is not None:
            t0 = (t1 - t0) * strength + t0
        assert t0 < t1, "ODE sampler has to be in forward time"

        def _sample(x, model, model_kwargs):
            return _euler_sample(x, model, model_kwargs, t0, t1, num_steps, do_shift, time_shifting_factor)

        return _sample


def _euler_sample(x, model, model_kwargs, t0, t1, num_steps, do_shift, time_shifting_factor):
    if not (x.is_cuda and x.dtype == th.bfloat16):
        raise ops._lib.VcbError("the sampler runs on CUDA bf16 latents only (visualcloze.py:399); there is no CPU path")
    B, Li, C = x.shape
    t = solver_grid(t0, t1, num_steps, Li, do_shift, time_shifting_factor)
    n_eval = int(num_steps) - 1
    kw = dict(model_kwargs)                       # the caller's dict is not mutated (transport.py:194-196 pops a copy)
    cond = kw.pop("cond", None)
    # FLUX time fed to the model: ones(B) * t -> 1 - t (integrators.py:109, transport.py:384), fp32.  torchdiffeq wraps the
    # ODE function in _PerturbFunc, which casts the evaluation time to the STATE dtype (t.to(y.abs().dtype)): with the bf16
    # latent the model sees 1 - bf16(tau_k) (up to ~2 units of 1000 t near tau = 1).  dt keeps the fp32 grid.
    t_vec = th.ones(n_eval, B) * t[:-1, None].to(x.dtype).float()
    flux_t = th.ones_like(t_vec) * (1 - t_vec)
    # dt is a 0-dim fp32 tensor multiplied into a bf16 tensor: it acts as a bf16 scalar
    dts = [float((t[k + 1] - t[k]).to(th.bfloat16)) for k in range(n_eval)]

    native = _is_native_forward(model)
    sp = model.__self__.engine()._sp if native else None
    if sp is not None:
        # single-image sequence parallelism: this rank integrates its own token rows (the ODE is row-local apart from the
        # attention inside the model); the trajectory is gathered once at the end.  The time grid above used the FULL Li.
        x = sp.shard(x).contiguous()
        cond = None if cond is None else sp.shard(cond)
    Li_full, Li = Li, x.shape[1]
    traj = th.empty(int(num_steps), B, Li, C, dtype=x.dtype, device=x.device)
    traj[0].copy_(x)
    Cc = 0 if cond is None else cond.shape[-1]
    inp = th.empty(B, Li, C + Cc, dtype=x.dtype, device=x.device)       # cat(x, cond) buffer (transport.py:195)
    ops.copy_cols(traj[0].reshape(B * Li, C), inp.reshape(B * Li, C + Cc), 0)
    if cond is not None:
        ops.copy_cols(cond.to(x.dtype).reshape(B * Li, Cc).contiguous(), inp.reshape(B * Li, C + Cc), C)

    if native:
        flux = model.__self__
        eng = flux.engine()
        if flux.params.guidance_embed and kw.get("guidance") is None:
            raise ValueError("Didn't get guidance strength for guidance distilled model.")
        eng.prepare(txt=kw["txt"], y=kw["y"], img_ids=kw["img_ids"], txt_ids=kw["txt_ids"], timesteps=flux_t,
                    guidance=kw.get("guidance"), txt_mask=kw.get("txt_mask"), img_mask=kw.get("img_mask"), n_img_tokens=Li_full)
        v = th.empty(B, Li, flux.params.out_channels, dtype=x.dtype, device=x.device)
    for k in range(n_eval):
        if native:
            eng.forward(k, inp, v)
            out = v
        else:
            out = model(inp, timesteps=flux_t[k].to(x.device), **kw)
        assert out.shape == x.shape, "Output shape from ODE solver must match input shape"
        ops.euler_update(traj[k].reshape(B * Li, C), out.reshape(B * Li, C), dts[k], traj[k + 1].reshape(B * Li, C),
                         inp.reshape(B * Li, C + Cc))
    if sp is not None:
        traj = sp.gather(traj, dim=2)
        sp.check()
    return traj
