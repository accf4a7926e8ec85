"""CPU restatement of the reference PGDM update (oracle; TEST INFRASTRUCTURE).
Follows samplers/samplers/pgdm.py:104-135:
    x0   = Tweedie(x_t)                                             networks/base.py:41-43
    loss = sum( (A^+ y - A^+ A x0)^2 )                              pgdm.py:116-121
    grad = d loss / d x_t                                           pgdm.py:123
    x    = ddim_step(x_t; e_t = x0) - guidance_weight*sqrt(1-acp_t)*grad        pgdm.py:125-135
"""
from __future__ import annotations

from typing import Callable, Sequence

import torch
from torch import Tensor

from .dps import bridge_coefficients, tweedie_x0
from .operators import OracleOperator


def pgdm_step(net, x_t: Tensor, *, t: int, t_prev: int, s: int, acp: Tensor, op: OracleOperator, y_flat: Tensor,
              guidance_weight: float, eta: float, z: Tensor) -> dict:
    x_t = x_t.detach().requires_grad_()
    x0 = tweedie_x0(x_t, net(x_t, t), acp[t])
    loss = (op.pinv(y_flat) - op.pinv(op.apply(x0))).pow(2).sum()
    (grad,) = torch.autograd.grad(loss, x_t)
    c_ell, c_s, std = bridge_coefficients(acp, t, t_prev, s, eta)
    xd = x_t.detach()
    ddim = (c_ell * xd + c_s * x0.detach()).to(xd.dtype) + std.to(xd.dtype) * z
    scale = guidance_weight * torch.sqrt(1 - acp[t])
    return {"x_next": (ddim - scale * grad).detach(), "grad": grad.detach(), "x0": x0.detach()}


def pgdm_sample(net, *, acp: Tensor, timesteps: Sequence[int], op: OracleOperator, y_flat: Tensor, leading: int,
                guidance_weight: float = 1.0, eta: float = 1.0, draw: Callable[[tuple], Tensor]) -> Tensor:
    ts = [int(v) for v in timesteps]
    x = draw((leading, *op.x_shape))
    for i in range(len(ts) - 1, 1, -1):
        x = pgdm_step(net, x, t=ts[i], t_prev=ts[i - 1], s=ts[0], acp=acp, op=op, y_flat=y_flat,
                      guidance_weight=guidance_weight, eta=eta, z=draw(tuple(x.shape)))["x_next"]
    with torch.no_grad():
        return tweedie_x0(x, net(x, ts[1]), acp[ts[1]])
