"""PSLDSampler -- Posterior Sampling with Latent Diffusion (Rout et al., 2023), with the call signature
and semantics of the reference sampler (samplers/samplers/psld.py:19-166).  One step:

    eps  = network(z_t, t)                       torch (graph kept)
    z0   = Tweedie(z_t, eps)                     psx_tweedie (autograd node)
    x0   = decode(z0)                            torch VAE decoder (differentiable)
    lik, x_eff = data term(x0)                   psx_dps_pre (r = y - A x0, |r|^2, A^T r) + psx_lincomb3
                 lik   = || y - A x0 ||_F        one norm over the whole local batch (psld.py:129-130)
                 x_eff = A^T y + x0 - A^T A x0   (psld.py:132-136; computed as x0 + A^T r)
    z_eff = encode(x_eff)                        torch VAE encoder (differentiable)
    glue = || z0 - z_eff ||_F                    psld.py:137-138
    grad = d(omega*lik + gamma*glue)/d z_t       torch autograd through UNet + VAE; the pixel-space block is
                                                 one autograd node whose backward is again psx_dps_pre
    z_{t-1} = c_ell z_t + c_s z0 + std*noise - grad        psx_bridge_update (psld.py:144-153)

The norms are batch-global exactly as in the reference.  On several GPUs each rank's shard
behaves like a separate reference run (rank-local norms) unless the sampler is built with ``process_group``:
then the two sums of squares are all-reduced (one scalar each per step) and the union of the shards reproduces
one reference call on the concatenated batch (samplers_b200/distributed.py).
"""
from __future__ import annotations

from typing import Callable, Generic, TypeVar

import torch
from torch import Tensor

from .. import _native
from ..distributed import global_norm, reduce_sum_
from ..inverse_problem import InverseProblem
from ..networks.base import LatentEpsilonNetwork, _TweedieFn, tweedie_scalars
from .base import PosteriorSampler
from .utils.batch_view import BatchView
from .utils.bridge_kernels import plan_steps

Condition_co = TypeVar("Condition_co", covariant=True)


class _PsldDataTerm(torch.autograd.Function):
    """(x0) -> (lik, x_eff) with  r = y - A x0,  lik = ||r||_F,  x_eff = x0 + A^T r.
    Backward: c_x0 = c_xeff - A^T A c_xeff - (c_lik / lik) * A^T r."""

    @staticmethod
    def forward(ctx, x0: Tensor, op, y: Tensor, obs_repeat: int, ws, zeros_y: Tensor, group=None):
        L, n = x0.shape
        atr = torch.empty_like(x0)
        part = torch.empty((L, op.err_parts), device=x0.device, dtype=torch.float32)
        # K1 with sa = 1, s1 = 0, w = 1:  "Tweedie" is the identity, cot = A^T r
        _native.dps_pre(op, x0, x0, y, obs_repeat, 1.0, 0.0, 1.0, atr, part, ws)
        lik = reduce_sum_(part.sum(), group).sqrt()   # batch-global over all ranks of `group`
        x_eff = torch.empty_like(x0)
        _native.lincomb3(x0, 1.0, atr, 1.0, None, 0.0, x_eff)
        ctx.op, ctx.ws, ctx.obs_repeat = op, ws, obs_repeat
        ctx.save_for_backward(atr, lik, zeros_y)
        return lik, x_eff

    @staticmethod
    def backward(ctx, c_lik: Tensor, c_xeff: Tensor):
        atr, lik, zeros_y = ctx.saved_tensors
        c = c_xeff.contiguous()
        L = c.shape[0]
        neg_ata_c = torch.empty_like(c)
        part = torch.empty((L, ctx.op.err_parts), device=c.device, dtype=torch.float32)
        _native.dps_pre(ctx.op, c, c, zeros_y, L, 1.0, 0.0, 1.0, neg_ata_c, part, ctx.ws)  # A^T(0 - A c)
        out = torch.empty_like(c)
        # -(c_lik / lik) * A^T r with both factors read on the device (no host synchronisation in the loop)
        _native.lincomb3_dev(c, 1.0, neg_ata_c, 1.0, atr, -1.0, c_lik.float().contiguous(), lik, out)
        return out, None, None, None, None, None, None


class PSLDSampler(PosteriorSampler, Generic[Condition_co]):
    draw: Callable = staticmethod(lambda shape, device, dtype: torch.randn(size=shape, device=device, dtype=dtype))

    def __init__(self, network, cuda_graph: bool = False, process_group=None):
        super().__init__(network, cuda_graph=False, process_group=process_group)
        if not isinstance(self._epsilon_network, LatentEpsilonNetwork):
            raise TypeError(
                f"{self.__class__.__name__} requires a latent diffusion model, but build_network returned a "
                f"non-latent network ({type(self._epsilon_network).__name__}).")

    def __call__(self, inverse_problem: InverseProblem, *, num_sampling_steps: int = 100,
                 num_reconstructions: int = 1, gamma: float = 1.0, omega: float = 0.1, eta: float = 1.0,
                 decode_output: bool = True, condition: Condition_co | None = None) -> Tensor:
        """Monte-Carlo reconstructions of shape (*batch_shape, num_reconstructions, *x_shape)
        (or the latents when ``decode_output`` is False)."""
        op = inverse_problem.operator
        x_shape = tuple(op.x_shape)
        x_view = BatchView(inverse_problem.batch_shape, num_reconstructions, x_shape)
        net: LatentEpsilonNetwork = self._epsilon_network
        latent_shape = tuple(net.get_latent_shape(x_shape))
        z_view = BatchView(inverse_problem.batch_shape, num_reconstructions, latent_shape)
        L = z_view.leading_size

        net.set_sampling_parameters(num_sampling_steps=num_sampling_steps, num_reconstructions=num_reconstructions,
                                    batch_size=x_view.batch_size)
        net.set_condition(condition)
        try:
            device, dtype = net.device, net.dtype
            if torch.device(device).type != "cuda":
                raise RuntimeError("PSLDSampler needs the network on a CUDA device (no CPU path)")
            # state and kernels are fp32; a half-precision network / VAE (scripts/run_psld.py:14 runs bf16) gets casts
            if dtype not in (torch.float32, torch.bfloat16, torch.float16):
                raise TypeError(f"PSLDSampler supports float32 / bfloat16 / float16 networks, got {dtype}")
            net_dtype, dtype = dtype, torch.float32

            def to_net(t: Tensor) -> Tensor:
                return t if net_dtype == torch.float32 else t.to(net_dtype)
            nat = op._native_cached(device)
            y = op._dense_observation(inverse_problem.observation.to(device=device, dtype=torch.float32))
            obs_repeat = num_reconstructions if x_view.batch_size > 1 else L
            wsb = nat.workspace_bytes(L)
            ws = torch.empty(wsb // 4, device=device, dtype=torch.float32) if wsb else None
            zeros_y = torch.zeros((1, nat.n_y), device=device, dtype=torch.float32)
            timesteps = [int(v) for v in net.timesteps.tolist()]
            plan = plan_steps(net.alphas_cumprod, timesteps, eta)

            z = self.draw((L, *latent_shape), device, dtype).contiguous()
            z_next = torch.empty_like(z)
            for sc in plan:
                z_in = z.detach().requires_grad_()
                eps = net.forward(to_net(z_in), sc.t).float()
                z0 = _TweedieFn.apply(z_in, eps, sc.sqrt_acp, sc.sqrt_1m_acp)
                x0 = net.decode(to_net(z0), differentiable=True).float()
                lik, x_eff = _PsldDataTerm.apply(x0.reshape(L, nat.n).contiguous(), nat, y, obs_repeat, ws, zeros_y,
                                                 self.process_group)
                z_eff = net.encode(to_net(x_eff.view(L, *x_shape)), differentiable=True).float()
                glue = global_norm(z0 - z_eff, self.process_group)
                (grad,) = torch.autograd.grad(omega * lik + gamma * glue, z_in)
                noise = self.draw(tuple(z.shape), device, dtype) if sc.std != 0.0 else None
                _native.bridge_update(z, eps.detach().contiguous(), noise, grad.contiguous(), sc.sqrt_acp,
                                      sc.sqrt_1m_acp, sc.c_ell, sc.c_s, sc.std, -1.0, z_next)
                z, z_next = z_next, z

            t1 = timesteps[1]
            sa, s1 = tweedie_scalars(net.alphas_cumprod, t1)
            with torch.no_grad():
                eps = net.forward(to_net(z), t1).float().contiguous()
                z0 = torch.empty_like(z)
                _native.tweedie(z.view(L, -1), eps.view(L, -1), sa, s1, z0.view(L, -1))
                if decode_output:
                    return x_view.unflatten(net.decode(to_net(z0), differentiable=False).float())
            return z_view.unflatten(z0)
        finally:
            net.clear_condition()
            net.clear_sampling_parameters()
