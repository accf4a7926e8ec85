"""CPU restatement of the reference ReSample update (oracle; TEST INFRASTRUCTURE).

Cited reference lines:
  eps-parameterised DDIM step   samplers/samplers/utils/bridge_kernels.py:82-115  (all in the acp dtype, no fp64)
  DPS conditioning              samplers/samplers/utils/resample_kernels.py:15-29, scale = acp_t * 0.5 (resample.py:145-148)
  pixel-space optimisation      resample_kernels.py:32-54   AdamW(lr 1e-2), MSE mean over ALL elements, stop loss < eps^2
  latent-space optimisation     resample_kernels.py:57-93   AdamW(lr 5e-3) through the decoder, plateau rule after iter 200
  stochastic resample           resample_kernels.py:96-107, sigma from :123-129
  control flow                  samplers/samplers/resample.py:113-225
"""
from __future__ import annotations

from typing import Callable, Sequence

import torch
from torch import Tensor

from .operators import OracleOperator


def ddim_eps_scalars(acp: Tensor, t: int, t_prev: int, eta: float) -> dict:
    """0-dim fp32 tensors exactly as bridge_kernels.py:95-111 computes them."""
    a_t, a_p = acp[t], acp[t_prev]
    sqrt_oma = (1 - a_t).sqrt()
    sigma_t = eta * ((1 - a_p) / (1 - a_t) * (1 - a_t / a_p)).clamp(min=0).sqrt()
    return {"sqrt_a_t": a_t.sqrt(), "sqrt_oma": sqrt_oma, "oma": 1 - a_t, "sqrt_a_p": a_p.sqrt(),
            "dir": (1 - a_p - sigma_t ** 2).clamp(min=0).sqrt(), "sigma_t": sigma_t}


def ddim_eps_step(x: Tensor, e: Tensor, acp: Tensor, t: int, t_prev: int, eta: float, noise: Tensor):
    s = ddim_eps_scalars(acp, t, t_prev, eta)
    pred_x0 = (x - s["sqrt_oma"] * e) / s["sqrt_a_t"]
    pseudo_x0 = (x - s["oma"] * e) / s["sqrt_a_t"]
    x_prev = s["sqrt_a_p"] * pred_x0 + s["dir"] * e + s["sigma_t"] * noise
    return x_prev, pred_x0, pseudo_x0


def compute_sigma(sigma_scale: float, a_t: Tensor, a_prev: Tensor) -> Tensor:
    return sigma_scale * (1 - a_prev) / (1 - a_t) * (1 - a_t / a_prev)


def stochastic_resample(pseudo_x0: Tensor, x_t: Tensor, a_t: Tensor, sigma: Tensor, noise: Tensor) -> Tensor:
    return (sigma * a_t.sqrt() * pseudo_x0 + (1 - a_t) * x_t) / (sigma + 1 - a_t) + \
        noise * torch.sqrt(1 / (1 / sigma + 1 / (1 - a_t)))


def pixel_optimization(y: Tensor, x_init: Tensor, op: OracleOperator, eps: float, max_iters: int):
    var = x_init.detach().clone().requires_grad_()
    opt = torch.optim.AdamW([var], lr=1e-2)
    iters = 0
    for _ in range(max_iters):
        opt.zero_grad()
        loss = torch.nn.functional.mse_loss(op.apply(var), y.expand_as(op.apply(var)) if y.shape != op.apply(var).shape else y)
        loss.backward()
        opt.step()
        iters += 1
        if loss.item() < eps ** 2:
            break
    return var.detach(), iters


def latent_optimization(y: Tensor, z_init: Tensor, op: OracleOperator, decode, eps: float, max_iters: int):
    z = z_init.detach().clone().requires_grad_()
    opt = torch.optim.AdamW([z], lr=5e-3)
    window: list[float] = []
    for itr in range(max_iters):
        opt.zero_grad()
        loss = torch.nn.functional.mse_loss(op.apply(decode(z)), y)
        loss.backward()
        opt.step()
        cur = loss.detach().item()
        if itr >= 200:
            window.append(cur)
            if len(window) > 1 and window[0] < cur:
                break
            if len(window) > 1:
                window.pop(0)
        if cur < eps ** 2:
            break
    return z.detach()


def resample_sample(eps_fn, decode, encode, *, acp: Tensor, timesteps: Sequence[int], op: OracleOperator,
                    y_flat: Tensor, latent_shape, leading: int, eps: float, sigma_scale: float = 40.0,
                    max_optimization_iters: int = 2000, eta: float = 1.0, inter_timesteps: int = 5,
                    time_travel_interval: int = 10, stage_splits: int = 3, draw: Callable[[tuple], Tensor],
                    decode_output: bool = True) -> Tensor:
    ts = [int(v) for v in timesteps]
    z = draw((leading, *latent_shape))
    total_steps = len(ts) - 1
    index_split = total_steps // stage_splits
    for idx in range(len(ts) - 1, 1, -1):
        t, tp = ts[idx], ts[idx - 1]
        with torch.no_grad():
            e = eps_fn(z, t)
        z_next, _, pseudo = ddim_eps_step(z.detach(), e, acp, t, tp, eta, draw(tuple(z.shape)))
        # DPS conditioning: gradient of the (batch-global) residual norm through the decoder only
        leaf = pseudo.detach().requires_grad_()
        norm = torch.linalg.norm(y_flat - op.apply(decode(leaf)))
        (g_pseudo,) = torch.autograd.grad(norm, leaf)
        norm_grad = g_pseudo / acp[t].sqrt()          # d pseudo / d z_t = 1 / sqrt(acp_t)   (eps is detached)
        z = z_next - norm_grad * (acp[t] * 0.5)
        if idx <= (total_steps - index_split) and idx > 0 and idx % time_travel_interval == 0:
            snapshot = z.detach().clone()
            for k in range(idx, max(idx - inter_timesteps, 1), -1):
                if k <= 1:
                    break
                with torch.no_grad():
                    e = eps_fn(z, ts[k])
                z, _, pseudo = ddim_eps_step(z.detach(), e, acp, ts[k], ts[k - 1], eta, draw(tuple(z.shape)))
            sigma = compute_sigma(sigma_scale, acp[t], acp[tp])
            if idx >= index_split:
                x_opt, _ = pixel_optimization(y_flat, decode(pseudo.detach()).detach(), op, eps, max_optimization_iters)
                z_opt = encode(x_opt).detach()
            else:
                z_opt = latent_optimization(y_flat, pseudo.detach(), op, decode, eps, max_optimization_iters)
            z = stochastic_resample(z_opt, snapshot, acp[tp], sigma, draw(tuple(z.shape)))
    z0 = latent_optimization(y_flat, z.detach(), op, decode, eps, max_optimization_iters)
    with torch.no_grad():
        return decode(z0) if decode_output else z0
