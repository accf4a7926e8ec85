"""CPU restatement of the reference PSLD update (oracle; TEST INFRASTRUCTURE).
Follows samplers/samplers/psld.py:111-163: per step
    z0   = Tweedie(z_t)                          networks/base.py:41-43
    x0   = decode(z0)                            psld.py:126
    lik  = || y_flat - A x0 ||_F   (ONE norm over the whole flat batch)      psld.py:129-130
    xeff = A^T y + x0 - A^T A x0                 psld.py:132-136
    glue = || z0 - encode(xeff) ||_F             psld.py:137-138
    grad = d(omega*lik + gamma*glue)/d z_t       psld.py:140-141
    z    = ddim_step(z_t; e_t = z0) - grad       psld.py:144-153
"""
from __future__ import annotations

from typing import Callable, Sequence

import torch
from torch import Tensor

from .dps import bridge_coefficients, tweedie_x0
from .operators import OracleOperator


def psld_step(eps_fn, decode, encode, z_t: Tensor, *, t: int, t_prev: int, s: int, acp: Tensor,
              op: OracleOperator, y_flat: Tensor, aty_flat: Tensor, omega: float, gamma: float, eta: float,
              noise: Tensor) -> dict:
    z_t = z_t.detach().requires_grad_()
    z0 = tweedie_x0(z_t, eps_fn(z_t, t), acp[t])
    x0 = decode(z0)
    hx0 = op.apply(x0)
    lik = torch.norm(y_flat - hx0)
    x_eff = aty_flat + x0 - op.adjoint(hx0)
    glue = torch.norm(z0 - encode(x_eff))
    (grad,) = torch.autograd.grad(omega * lik + gamma * glue, z_t)
    c_ell, c_s, std = bridge_coefficients(acp, t, t_prev, s, eta)
    zd = z_t.detach()
    mean = (c_ell * zd + c_s * z0.detach()).to(zd.dtype)
    z_next = mean + std.to(zd.dtype) * noise - grad
    return {"z_next": z_next.detach(), "z0": z0.detach(), "grad": grad.detach(), "lik": lik.detach(),
            "glue": glue.detach()}


def psld_sample(eps_fn, decode, encode, *, acp: Tensor, timesteps: Sequence[int], op: OracleOperator,
                y_flat: Tensor, latent_shape, leading: int, omega: float = 0.1, gamma: float = 1.0,
                eta: float = 1.0, draw: Callable[[tuple], Tensor], decode_output: bool = True) -> Tensor:
    ts = [int(v) for v in timesteps]
    aty = op.adjoint(y_flat)
    z = draw((leading, *latent_shape))
    for i in range(len(ts) - 1, 1, -1):
        noise = draw(tuple(z.shape))
        z = psld_step(eps_fn, decode, encode, z, t=ts[i], t_prev=ts[i - 1], s=ts[0], acp=acp, op=op,
                      y_flat=y_flat, aty_flat=aty, omega=omega, gamma=gamma, eta=eta, noise=noise)["z_next"]
    with torch.no_grad():
        z0 = tweedie_x0(z, eps_fn(z, ts[1]), acp[ts[1]])
        return decode(z0) if decode_output else z0
