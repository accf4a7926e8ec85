"""CPU restatement of the reference DPS update (oracle; TEST INFRASTRUCTURE).

Two forms of one timestep are provided:

* ``dps_step_autograd``  -- the literal algorithm: Tweedie estimate, summed
  log-likelihood, ``torch.autograd.grad`` back to the noisy state, bridge step,
  norm-scaled guidance.  Follows samplers/samplers/dps.py:91-122.
* ``dps_step_closed_form`` -- the same update with the likelihood gradient
  written out as an explicit cotangent + one network VJP, i.e. the shape the
  CUDA kernels K1/K2 compute.  Checked against the
  autograd form and against recordings of the unmodified reference
  (tests/test_oracle_golden.py).

Cited reference lines:
  Tweedie            samplers/networks/base.py:41-43
  residual           samplers/inverse_problem.py:17-18
  Gaussian log-prob  samplers/noise.py:77-79
  Poisson  log-prob  samplers/noise.py:121-123
  bridge statistics  samplers/samplers/utils/bridge_kernels.py:15-46
  ancestral draw     samplers/samplers/utils/bridge_kernels.py:49-59, 62-75
  guidance           samplers/samplers/dps.py:116-122
  final estimate     samplers/samplers/dps.py:125-130
"""
from __future__ import annotations

from typing import Callable, Sequence

import torch
from torch import Tensor

from .operators import OracleOperator

NetFn = Callable[[Tensor, int], Tensor]  # (x_t, t) -> eps, differentiable in x_t


# ----------------------------------------------------------------------------
# scalar pieces
# ----------------------------------------------------------------------------
def tweedie_x0(x_t: Tensor, eps: Tensor, acp_t: Tensor) -> Tensor:
    """networks/base.py:42-43 -- three separately rounded tensor ops."""
    return (x_t - (1 - acp_t) ** 0.5 * eps) / (acp_t ** 0.5)


def log_prob(residual: Tensor, noise_kind: str, noise_param: Tensor) -> Tensor:
    dims = tuple(range(1, residual.ndim))
    if noise_kind == "gaussian":  # noise.py:78-79, noise_param = sigma
        var = noise_param.pow(2)
        return -(residual.square().sum(dim=dims)) / (2 * var)
    if noise_kind == "poisson":  # noise.py:123, noise_param = rate
        return -(residual.pow(2) / (noise_param + 1e-3)).sum(dim=dims)
    raise ValueError(noise_kind)


def likelihood_weight(noise_kind: str, noise_param: Tensor) -> float:
    """w such that d(sum log_prob)/d(residual) = -w * residual (fp64 host value)."""
    p = float(noise_param)
    if noise_kind == "gaussian":
        return 1.0 / (float(torch.tensor(p, dtype=torch.float32).pow(2)))
    if noise_kind == "poisson":
        return 2.0 / float(torch.tensor(p, dtype=torch.float32) + 1e-3)
    raise ValueError(noise_kind)


def bridge_coefficients(acp: Tensor, ell: int, t: int, s: int, eta: float):
    """bridge_kernels.py:28-39 in fp64; returns 0-dim fp64 (c_ell, c_s, std)."""
    a_t = acp[t].to(torch.float64)
    a_ell = acp[ell].to(torch.float64)
    a_s = acp[s].to(torch.float64)
    s_to_t = a_t / a_s
    t_to_ell = a_ell / a_t
    s_to_ell = a_ell / a_s
    std = eta * ((1 - t_to_ell) * (1 - s_to_t) / (1 - s_to_ell)) ** 0.5
    c_ell = ((1 - s_to_t - std ** 2) / (1 - s_to_ell)) ** 0.5
    c_s = (s_to_t ** 0.5) - c_ell * (s_to_ell ** 0.5)
    return c_ell, c_s, std


# ----------------------------------------------------------------------------
# one timestep, literal (autograd) form
# ----------------------------------------------------------------------------
def dps_step_autograd(net: NetFn, x_t: Tensor, *, t: int, t_prev: int, s: int,
                      acp: Tensor, op: OracleOperator, y: Tensor,
                      noise_kind: str, noise_param: Tensor,
                      gamma: float, eta: float, z: Tensor) -> dict:
    L = x_t.shape[0]
    x_t = x_t.detach().requires_grad_()
    eps = net(x_t, t)
    x0 = tweedie_x0(x_t, eps, acp[t])
    ll = log_prob(y - op.apply(x0), noise_kind, noise_param).sum()
    grad = torch.autograd.grad(ll, x_t)[0]

    c_ell, c_s, std = bridge_coefficients(acp, t, t_prev, s, eta)
    x_d = x_t.detach()
    mean = (c_ell * x_d + c_s * x0.detach()).to(x_d.dtype)
    x_next = mean + std.to(x_d.dtype) * z
    with torch.no_grad():
        r = (y - op.apply(x0)).reshape(L, -1)
        err = r.norm(dim=1).view(L, *([1] * (x_d.ndim - 1)))
        x_next = x_next + (gamma / (err + 1e-9)) * grad
    return {"x_next": x_next.detach(), "x0": x0.detach(), "eps": eps.detach(),
            "grad": grad.detach(), "err": err.reshape(L).detach()}


# ----------------------------------------------------------------------------
# one timestep, closed form (what K1 / K2 compute)
# ----------------------------------------------------------------------------
def k1_reference(x_t: Tensor, eps: Tensor, *, acp_t: Tensor, op: OracleOperator,
                 y: Tensor, weight: float):
    """Returns (d, e2, x0): d = w * A^T(y - A x0) / sqrt(acp_t), e2_i = |r_i|^2."""
    L = x_t.shape[0]
    sa = acp_t ** 0.5
    x0 = tweedie_x0(x_t, eps, acp_t)
    r = y - op.apply(x0)
    e2 = r.reshape(L, -1).square().sum(dim=1)
    g0 = op.adjoint(r.expand(L, *r.shape[1:]) if r.shape[0] != L else r) * weight
    d = g0 / sa
    return d, e2, x0


def k2_reference(x_t: Tensor, eps: Tensor, d: Tensor, v: Tensor, z: Tensor | None,
                 e2: Tensor, *, acp_t: Tensor, c_ell: Tensor, c_s: Tensor,
                 std: Tensor, gamma: float):
    """x_next = c_ell x_t + c_s x0 + std z + gamma / (sqrt(e2)+1e-9) * (d - s1 v)."""
    L = x_t.shape[0]
    s1 = (1 - acp_t) ** 0.5
    x0 = tweedie_x0(x_t, eps, acp_t)
    mean = (c_ell * x_t + c_s * x0).to(x_t.dtype)
    x_next = mean if z is None else mean + std.to(x_t.dtype) * z
    grad = d + (-s1) * v
    err = e2.sqrt().view(L, *([1] * (x_t.ndim - 1)))
    return x_next + (gamma / (err + 1e-9)) * grad, grad


def dps_step_closed_form(net: NetFn, x_t: Tensor, *, t: int, t_prev: int, s: int,
                         acp: Tensor, op: OracleOperator, y: Tensor,
                         noise_kind: str, noise_param: Tensor,
                         gamma: float, eta: float, z: Tensor) -> dict:
    x_t = x_t.detach().requires_grad_()
    eps = net(x_t, t)
    w = likelihood_weight(noise_kind, noise_param)
    with torch.no_grad():
        d, e2, x0 = k1_reference(x_t.detach(), eps.detach(), acp_t=acp[t], op=op, y=y, weight=w)
    (v,) = torch.autograd.grad(eps, x_t, grad_outputs=d)
    c_ell, c_s, std = bridge_coefficients(acp, t, t_prev, s, eta)
    with torch.no_grad():
        x_next, grad = k2_reference(x_t.detach(), eps.detach(), d, v, z, e2,
                                    acp_t=acp[t], c_ell=c_ell, c_s=c_s, std=std, gamma=gamma)
    return {"x_next": x_next, "x0": x0, "eps": eps.detach(), "grad": grad,
            "err": e2.sqrt(), "d": d, "v": v, "e2": e2}


# ----------------------------------------------------------------------------
# the whole sampler (dps.py:83-130), with the noise draws supplied by the caller
# ----------------------------------------------------------------------------
def dps_sample(net: NetFn, *, acp: Tensor, timesteps: Sequence[int],
               op: OracleOperator, y: Tensor, noise_kind: str, noise_param: Tensor,
               leading: int, gamma: float = 1.0, eta: float = 1.0,
               draw: Callable[[tuple], Tensor], form: str = "autograd",
               record: list | None = None) -> Tensor:
    ts = [int(v) for v in timesteps]
    step = dps_step_autograd if form == "autograd" else dps_step_closed_form
    x = draw((leading, *op.x_shape))
    for i in range(len(ts) - 1, 1, -1):
        z = draw(tuple(x.shape))
        out = step(net, x, t=ts[i], t_prev=ts[i - 1], s=ts[0], acp=acp, op=op, y=y,
                   noise_kind=noise_kind, noise_param=noise_param,
                   gamma=gamma, eta=eta, z=z)
        if record is not None:
            record.append({"t": ts[i], "t_prev": ts[i - 1], "x_t": x.detach().clone(),
                           "z": z, **out})
        x = out["x_next"]
    with torch.no_grad():
        eps = net(x, ts[1])
        return tweedie_x0(x, eps, acp[ts[1]])


def psnr(a: Tensor, b: Tensor) -> float:
    """scripts/run_resample.py:30-34 -- data in [-1, 1], peak-to-peak 2."""
    mse = torch.mean((a - b) ** 2)
    return float(10 * torch.log10(4.0 / mse))
