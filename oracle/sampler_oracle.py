"""CPU oracle for the flow-matching Euler sampler -- TEST INFRASTRUCTURE ONLY.

Restates:
  * ``transport/transport.py:361-410``   Sampler.sample_ode (time interval, SDEdit strength,
                                           FLUX time flip ``1 - t`` and output negation :384)
  * ``transport/integrators.py:79-120``  ode.__init__ / ode.sample (linspace grid, time-shifting
                                           factor, resolution-dependent shift)
  * ``transport/utils.py:33-43``         time_shift (mirrored form), get_lin_function
  * ``transport/transport.py:193-198``   velocity_ode (``cat(x, cond)`` then the model call)
  * ``torchdiffeq.odeint(method="euler")`` -- third-party, unpinned, absent from /root/reference:
    fixed-grid Euler, ``y[k+1] = y[k] + (t[k+1] - t[k]) * f(t[k], y[k])``, returns the stacked
    trajectory.  With a bf16 state and a 0-dim fp32 ``dt`` tensor, torch type promotion makes
    the update ``bf16(y + bf16(bf16(dt) * f))`` (probed).  ``odeint`` wraps ``f`` in
    ``_PerturbFunc``, whose forward casts the evaluation time to the state dtype (``t.to(y.abs().dtype)``), so
    a bf16 state makes the model see ``bf16(t[k])``; ``dt`` is taken from the un-rounded fp32 grid.

Parity status: pinned against the reference ``transport`` package run in the build container
with a 12-line Euler ``torchdiffeq`` shim (``oracle/gen_golden.py``).
"""
from __future__ import annotations

import math

import torch


def solver_grid(num_steps: int, seq_len: int, do_shift: bool = True, time_shifting_factor=None,
                strength=None) -> torch.Tensor:
    """The solver's time grid tau (0 = noise -> 1 = data), fp32 [num_steps]."""
    t0, t1 = 0.0, 1.0                                    # check_interval, LINEAR+VELOCITY
    if strength is not None:
        t0 = (t1 - t0) * strength + t0                   # transport.py:395-396
    t = torch.linspace(t0, t1, num_steps)                # integrators.py:99
    if time_shifting_factor:
        t = t / (t + time_shifting_factor - time_shifting_factor * t)   # :100-101
    if do_shift:
        m = (1.15 - 0.5) / (4096 - 256)                  # utils.py:40-43, y1=.5 y2=1.15
        mu = m * seq_len + (0.5 - m * 256)
        tt = 1 - t                                       # utils.py:33-38 (sigma = 1)
        tt = math.exp(mu) / (math.exp(mu) + (1 / tt - 1) ** 1.0)
        t = 1 - tt
    return t


def sample_ode(x: torch.Tensor, model_fn, model_kwargs: dict, num_steps: int, do_shift: bool = True,
               time_shifting_factor=None, strength=None) -> torch.Tensor:
    """Returns the whole trajectory [num_steps, B, Li, C] like the reference sampler."""
    t = solver_grid(num_steps, x.shape[1], do_shift, time_shifting_factor, strength).to(x.device)   # integrators.py:112
    kw = dict(model_kwargs)
    cond = kw.pop("cond", None)
    ys = [x]
    y = x
    for k in range(num_steps - 1):
        tau = t[k].to(y.dtype)                                     # torchdiffeq _PerturbFunc: t.to(y.abs().dtype)
        t_vec = torch.ones(y.shape[0]).to(y.device) * tau          # integrators.py:109
        t_flux = torch.ones_like(t_vec) * (1 - t_vec)              # transport.py:384
        inp = y if cond is None else torch.cat((y, cond), dim=-1)  # transport.py:194-196
        f = -model_fn(inp, timesteps=t_flux, **kw)
        dt = t[k + 1] - t[k]                                       # 0-dim fp32 tensor
        y = y + dt * f                                             # promotes to y.dtype
        ys.append(y)
    return torch.stack(ys, 0)
