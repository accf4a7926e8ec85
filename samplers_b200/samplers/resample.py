"""ReSampleSampler -- Solving Inverse Problems with Latent Diffusion Models via Hard Data Consistency
(Song et al., 2023), with the call signature and control flow of the reference sampler
(samplers/samplers/resample.py:26-228) and its arithmetic in sm_100a kernels:

  eps-DDIM step            psx_ddim_eps_step            bridge_kernels.py:82-115
  DPS conditioning         residual-norm autograd node built on psx_dps_pre (r, |r|^2, A^T r) whose gradient
                           flows through the VAE decoder only; z <- z_next - (acp_t/2) * grad
                           (resample_kernels.py:15-29, resample.py:145-158)
  pixel-space optimisation on-device loop: psx_dps_pre (gradient of the batch-mean MSE) + psx_adamw_step
                           with a device-side stop flag -- one host sync per CHECK_EVERY iterations
                           instead of one per iteration (resample_kernels.py:32-54)
  latent-space optimisation psx_adamw_step on the latent, decoder forward/backward in torch, the reference's
                           plateau / eps^2 stopping rules on the host (resample_kernels.py:57-93)
  stochastic resample      psx_stochastic_resample      resample_kernels.py:96-107, :123-129

As in the reference the residual norm / MSE are batch-global, and ``scale`` is accepted
but unused (App. B-9).  With ``process_group`` the sums behind them are all-reduced over the ranks (one scalar per
reduction), so equal shards of a batch reproduce one reference call on the whole batch, early stops included.
"""
from __future__ import annotations

from typing import Callable, Generic, TypeVar

import torch
from torch import Tensor

from .. import _native
from ..distributed import group_size, reduce_sum_
from ..inverse_problem import InverseProblem
from ..networks.base import LatentEpsilonNetwork
from ..noise import GaussianNoise
from .base import PosteriorSampler
from .utils.batch_view import BatchView

Condition_co = TypeVar("Condition_co", covariant=True)

CHECK_EVERY = 32  # pixel-space optimiser: iterations between reads of the device stop flag


def ddim_eps_scalars(acp: Tensor, t: int, t_prev: int, eta: float) -> dict:
    """The six fp32 scalars of bridge_kernels.py:95-111, evaluated with the same 0-dim tensor ops."""
    a = acp.detach().to(device="cpu", dtype=torch.float32)
    a_t, a_p = a[int(t)], a[int(t_prev)]
    sigma_t = eta * ((1 - a_p) / (1 - a_t) * (1 - a_t / a_p)).clamp(min=0).sqrt()
    return {"sqrt_a_t": float(a_t.sqrt()), "sqrt_oma": float((1 - a_t).sqrt()), "oma": float(1 - a_t),
            "sqrt_a_p": float(a_p.sqrt()), "dir": float((1 - a_p - sigma_t ** 2).clamp(min=0).sqrt()),
            "sigma_t": float(sigma_t)}


def resample_scalars(acp: Tensor, t: int, t_prev: int, sigma_scale: float) -> tuple[float, float, float, float]:
    """(c_p, c_x, den, k_n) of the stochastic resample at a = acp[t_prev], sigma from (acp[t], acp[t_prev])."""
    a = acp.detach().to(device="cpu", dtype=torch.float32)
    a_t, a_p = a[int(t)], a[int(t_prev)]
    sigma = sigma_scale * (1 - a_p) / (1 - a_t) * (1 - a_t / a_p)
    return (float(sigma * a_p.sqrt()), float(1 - a_p), float(sigma + 1 - a_p),
            float(torch.sqrt(1 / (1 / sigma + 1 / (1 - a_p)))))


class _ResidualTerm(torch.autograd.Function):
    """x -> ||y - A x||_F (mode 'norm') or mean((y - A x)^2) over all elements (mode 'mse'); one K1 launch
    forward (partials + A^T r), one scaling backward."""

    @staticmethod
    def forward(ctx, x: Tensor, op, y: Tensor, obs_repeat: int, ws, mode: str, group=None, n_obs=None):
        # n_obs: observed values per sample as the reference's nn.MSELoss counts them (prod(operator.y_shape): the m
        # kept pixels of a flattened inpainting operator, not the n entries of the dense-mask form the kernels use)
        L = x.shape[0]
        atr = torch.empty_like(x)
        part = torch.empty((L, op.err_parts), device=x.device, dtype=torch.float32)
        _native.dps_pre(op, x, x, y, obs_repeat, 1.0, 0.0, 1.0, atr, part, ws)
        e2 = reduce_sum_(part.sum(), group)             # batch-global over all ranks of `group`
        count = float(L * (op.n_y if n_obs is None else n_obs)) * group_size(group)   # equal shards
        val = e2.sqrt() if mode == "norm" else e2 / count
        ctx.mode, ctx.count = mode, count
        ctx.save_for_backward(atr, val)
        return val

    @staticmethod
    def backward(ctx, c: Tensor):
        atr, val = ctx.saved_tensors
        out = torch.empty_like(atr)
        c = c.float().contiguous()
        # kappa * A^T r, kappa = -(c / val) or -(2 / count) * c, completed on the device (no host synchronisation)
        if ctx.mode == "norm":
            _native.lincomb3_dev(atr, 0.0, atr, 0.0, atr, -1.0, c, val, out)
        else:
            _native.lincomb3_dev(atr, 0.0, atr, 0.0, atr, -2.0 / ctx.count, c, None, out)
        return out, None, None, None, None, None, None, None


class ReSampleSampler(PosteriorSampler, Generic[Condition_co]):
    draw: Callable = staticmethod(lambda shape, device, dtype: torch.randn(size=shape, device=device, dtype=dtype))

    def __init__(self, network, cuda_graph: bool = False, process_group=None):
        super().__init__(network, cuda_graph=False, process_group=process_group)
        if not isinstance(self._epsilon_network, LatentEpsilonNetwork):
            raise TypeError(
                f"{self.__class__.__name__} requires a latent diffusion model, but received a non-latent network "
                f"({type(self._epsilon_network).__name__}).")

    # ------------------------------------------------------------------ hard data consistency
    def _pixel_optimization(self, nat, y, obs_repeat, ws, x_init: Tensor, eps: float, max_iters: int,
                            n_obs=None) -> Tensor:
        L, n = x_init.shape
        x = x_init.detach().clone()
        m, v, grad = torch.zeros_like(x), torch.zeros_like(x), torch.empty_like(x)
        part = torch.empty((L, nat.err_parts), device=x.device, dtype=torch.float32)
        flags = torch.zeros(2, device=x.device, dtype=torch.int32)
        group = getattr(self, "process_group", None)
        count = float(L * (nat.n_y if n_obs is None else n_obs)) * group_size(group)   # resample_kernels.py:48 (MSELoss)
        for i in range(max_iters):
            # grad of mean((y - A x)^2) = -(2/N) A^T (y - A x); partials of |r|^2 in the same launch
            _native.dps_pre(nat, x, x, y, obs_repeat, 1.0, 0.0, -2.0 / count, grad, part, ws)
            # several ranks: the stop rule looks at the GLOBAL mean -> one scalar all-reduce per iteration, still
            # without a host sync (the kernel reads the reduced sum from device memory)
            loss_parts = part if group_size(group) == 1 else reduce_sum_(part.sum().reshape(1), group)
            _native.adamw_step(x, grad, m, v, 1e-2, i + 1, flags=flags, flag_in=i & 1, loss_parts=loss_parts,
                               loss_scale=1.0 / count, loss_threshold=eps ** 2)
            if (i + 1) % CHECK_EVERY == 0 and int(flags[(i + 1) & 1]) != 0:
                break
        return x

    def _latent_optimization(self, net, nat, y, obs_repeat, ws, z_init: Tensor, x_shape, eps: float,
                             max_iters: int, n_obs=None) -> Tensor:
        z = z_init.detach().clone()
        m, v = torch.zeros_like(z), torch.zeros_like(z)
        L = z.shape[0]
        window: list[float] = []
        for itr in range(max_iters):
            leaf = z.detach().requires_grad_()
            nd = getattr(self, "_net_dtype", torch.float32)
            x = net.decode(leaf if nd == torch.float32 else leaf.to(nd), differentiable=True).float()
            loss = _ResidualTerm.apply(x.reshape(L, nat.n).contiguous(), nat, y, obs_repeat, ws, "mse",
                                       self.process_group, n_obs)
            (g,) = torch.autograd.grad(loss, leaf)
            _native.adamw_step(z, g.contiguous(), m, v, 5e-3, itr + 1)
            cur = float(loss.detach())
            if itr >= 200:
                window.append(cur)
                if len(window) > 1 and window[0] < cur:
                    break
                if len(window) > 1:
                    window.pop(0)
            if cur < eps ** 2:
                break
        return z

    # ------------------------------------------------------------------ the sampler
    def __call__(self, inverse_problem: InverseProblem, *, num_sampling_steps: int = 100,
                 num_reconstructions: int = 1, scale: float = 0.3, sigma_scale: float = 40.0,
                 max_optimization_iters: int = 2000, eta: float = 1.0, inter_timesteps: int = 5,
                 time_travel_interval: int = 10, stage_splits: int = 3, decode_output: bool = True,
                 condition: Condition_co | None = None) -> Tensor:
        """Reconstructions of shape (*batch_shape, num_reconstructions, *x_shape) (latents if not decoded)."""
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
                raise RuntimeError("ReSampleSampler needs the network on a CUDA device (no CPU path)")
            # state and kernels are fp32; a half-precision network / VAE gets casts at its boundary
            if dtype not in (torch.float32, torch.bfloat16, torch.float16):
                raise TypeError(f"ReSampleSampler supports float32 / bfloat16 / float16 networks, got {dtype}")
            net_dtype, dtype = dtype, torch.float32
            self._net_dtype = net_dtype

            def to_net(t: Tensor) -> Tensor:
                return t if net_dtype == torch.float32 else t.to(net_dtype)
            nat = op._native_cached(device)
            n_obs = 1
            for d in op.y_shape:
                n_obs *= int(d)
            y = op._dense_observation(inverse_problem.observation.to(device=device, dtype=torch.float32))
            obs_repeat = num_reconstructions if x_view.batch_size > 1 else L
            wsb = nat.workspace_bytes(L)
            ws = torch.empty(wsb // 4, device=device, dtype=torch.float32) if wsb else None
            acp = net.alphas_cumprod
            ts = [int(v) for v in net.timesteps.tolist()]
            eps = float(inverse_problem.noise.sigma) if isinstance(inverse_problem.noise, GaussianNoise) else 1e-3
            total_steps = len(ts) - 1
            index_split = total_steps // stage_splits
            acp_host = acp.detach().to(device="cpu", dtype=torch.float32)

            def eps_ddim(z_cur: Tensor, t: int, t_prev: int):
                with torch.no_grad():
                    e = net.predict_noise(to_net(z_cur), t).float().contiguous()
                sc = ddim_eps_scalars(acp, t, t_prev, eta)
                noise = self.draw(tuple(z_cur.shape), device, dtype) if sc["sigma_t"] != 0.0 else None
                z_prev, pseudo = torch.empty_like(z_cur), torch.empty_like(z_cur)
                _native.ddim_eps_step(z_cur, e, noise, sc, z_prev, None, pseudo)
                return z_prev, pseudo

            z = self.draw(z_view.flat_shape, device, dtype).contiguous()
            for idx in range(len(ts) - 1, 1, -1):
                t, tp = ts[idx], ts[idx - 1]
                z_next, pseudo = eps_ddim(z, t, tp)
                # DPS conditioning: d||y - A D(pseudo)|| / d z_t, with d pseudo / d z_t = 1/sqrt(acp_t)
                leaf = pseudo.requires_grad_()
                x = net.decode(to_net(leaf), differentiable=True).float()
                norm = _ResidualTerm.apply(x.reshape(L, nat.n).contiguous(), nat, y, obs_repeat, ws, "norm",
                                           self.process_group)
                (g_pseudo,) = torch.autograd.grad(norm, leaf)
                a_t = acp_host[t]
                z = torch.empty_like(z_next)
                _native.lincomb3(z_next, 1.0, g_pseudo.contiguous(), -float((a_t * 0.5) / a_t.sqrt()), None, 0.0, z)
                pseudo = pseudo.detach()

                if idx <= (total_steps - index_split) and idx > 0 and idx % time_travel_interval == 0:
                    snapshot = z.clone()
                    for k in range(idx, max(idx - inter_timesteps, 1), -1):
                        if k <= 1:
                            break
                        z, pseudo = eps_ddim(z, ts[k], ts[k - 1])
                    c_p, c_x, den, k_n = resample_scalars(acp, t, tp, sigma_scale)
                    if idx >= index_split:
                        x_pix = net.decode(to_net(pseudo), differentiable=False).float().reshape(L, nat.n).contiguous()
                        x_opt = self._pixel_optimization(nat, y, obs_repeat, ws, x_pix, eps, max_optimization_iters, n_obs)
                        z_opt = net.encode(to_net(x_opt.view(L, *x_shape)), differentiable=False).float().contiguous()
                    else:
                        z_opt = self._latent_optimization(net, nat, y, obs_repeat, ws, pseudo, x_shape, eps,
                                                          max_optimization_iters, n_obs)
                    noise = self.draw(tuple(z.shape), device, dtype)
                    z = torch.empty_like(snapshot)
                    _native.stochastic_resample(z_opt, snapshot, noise, c_p, c_x, den, k_n, z)

            z0 = self._latent_optimization(net, nat, y, obs_repeat, ws, z, x_shape, eps, max_optimization_iters, n_obs)
            if decode_output:
                return x_view.unflatten(net.decode(to_net(z0), differentiable=False).float())
            return z_view.unflatten(z0)
        finally:
            net.clear_condition()
            net.clear_sampling_parameters()
