"""Host-side step planner for the bridge (DDIM / DDPM ancestral) update.

The reference recomputes the bridge statistics on the device every step as ~25
0-dim fp64 launches plus an implicit sync (samplers/samplers/utils/
bridge_kernels.py:15-46, :70).  Here the whole schedule is turned into a table
of fp32 scalars ONCE per call, on the host, with the reference's dtype path:

  Tweedie scalars  sqrt(acp_t), sqrt(1-acp_t): fp32 ops on the fp32 table
                   (networks/base.py:42-43)
  bridge scalars   c_ell, c_s, std: fp64 from the fp32 table entries
                   (bridge_kernels.py:28-39), rounded to fp32 when applied
                   (a 0-dim fp64 tensor times an fp32 tensor is an fp32 op with
                   the coefficient rounded to fp32 -- SURVEY section 5 probe)

``ddim_step`` / ``sample_bridge_kernel`` / ``compute_bridge_kernel_statistics``
keep the reference's names and argument meaning for users of the utilities.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Sequence

import torch
from torch import Tensor


@dataclass(frozen=True)
class BridgeStatistics:
    mean: Tensor
    std: Tensor


@dataclass(frozen=True)
class StepScalars:
    """Everything one guided timestep needs, as Python floats holding fp32 values."""
    t: int
    t_prev: int
    sqrt_acp: float       # sqrt(acp[t])
    sqrt_1m_acp: float    # sqrt(1 - acp[t])
    c_ell: float
    c_s: float
    std: float


def bridge_coefficients(acp: Tensor, ell: int, t: int, s: int, eta: float) -> tuple[Tensor, Tensor, Tensor]:
    """(c_ell, c_s, std) as 0-dim fp64 tensors; s < t < ell index the padded acp table."""
    a = acp.detach().to("cpu")
    a_t, a_ell, a_s = (a[int(i)].to(torch.float64) for i in (t, ell, s))
    r_st, r_tl, r_sl = a_t / a_s, a_ell / a_t, a_ell / a_s
    std = eta * ((1 - r_tl) * (1 - r_st) / (1 - r_sl)) ** 0.5
    c_ell = ((1 - r_st - std ** 2) / (1 - r_sl)) ** 0.5
    c_s = r_st ** 0.5 - c_ell * r_sl ** 0.5
    return c_ell, c_s, std


def _f32(v: Tensor) -> float:
    return float(v.to(torch.float32))


def plan_steps(acp: Tensor, timesteps: Sequence[int], eta: float) -> list[StepScalars]:
    """Scalar table for the guided iterations i = len(ts)-1 .. 2 (dps.py:91-93)."""
    a = acp.detach().to(device="cpu", dtype=torch.float32)
    ts = [int(v) for v in timesteps]
    plan = []
    for i in range(len(ts) - 1, 1, -1):
        t, p = ts[i], ts[i - 1]
        c_ell, c_s, std = bridge_coefficients(a, t, p, ts[0], eta)
        plan.append(StepScalars(t=t, t_prev=p, sqrt_acp=float(a[t] ** 0.5), sqrt_1m_acp=float((1 - a[t]) ** 0.5),
                                c_ell=_f32(c_ell), c_s=_f32(c_s), std=_f32(std)))
    return plan


STEP_ROW = 8  # PSX_STEP_ROW (include/psx.h)


def step_rows(plan: Sequence[StepScalars], weight: float, gamma) -> Tensor:
    """Device-table form of ``plan`` for psx_dps_pre_dev / psx_dps_post_dev: one fp32 row
    [sqrt_acp, sqrt_1m_acp, weight / sqrt_acp, c_ell, c_s, std, gamma, 0] per guided step, holding exactly the
    fp32 values the by-value entry points receive (weight / sqrt_acp is one fp32 division, as inside libpsx).
    ``gamma`` is a float or a callable ``StepScalars -> float`` (PGDM's per-step scale)."""
    rows = torch.zeros((len(plan), STEP_ROW), dtype=torch.float32)
    w = torch.tensor(float(weight), dtype=torch.float32)
    for k, sc in enumerate(plan):
        g = gamma(sc) if callable(gamma) else gamma
        coef = float(w / torch.tensor(sc.sqrt_acp, dtype=torch.float32))
        rows[k, :7] = torch.tensor([sc.sqrt_acp, sc.sqrt_1m_acp, coef, sc.c_ell, sc.c_s, sc.std, float(g)],
                                   dtype=torch.float64).to(torch.float32)
    return rows


# ---------------------------------------------------------------------------- utilities with the reference's names
def compute_bridge_kernel_statistics(x_ell: Tensor, x_s: Tensor, epsilon_net, ell: int, t: int, s: int,
                                     eta: float = 1.0) -> BridgeStatistics:
    c_ell, c_s, std = bridge_coefficients(epsilon_net.alphas_cumprod, ell, t, s, eta)
    mean = _f32(c_ell) * x_ell + _f32(c_s) * x_s
    return BridgeStatistics(mean=mean, std=torch.tensor(_f32(std), dtype=x_ell.dtype, device=x_ell.device))


def sample_bridge_kernel(x_ell: Tensor, x_s: Tensor, epsilon_net, ell: int, t: int, s: int, eta: float = 1.0):
    st = compute_bridge_kernel_statistics(x_ell, x_s, epsilon_net, ell, t, s, eta)
    return st.mean + st.std * torch.randn_like(st.mean)


def ddim_step(x: Tensor, epsilon_net, t: float, t_prev: float, eta: float, e_t: Tensor | None = None):
    if e_t is None:
        e_t = epsilon_net.predict_x0(x, t)
    return sample_bridge_kernel(x_ell=x, x_s=e_t, epsilon_net=epsilon_net, ell=int(t), t=int(t_prev),
                                s=int(epsilon_net.timesteps[0]), eta=eta)
