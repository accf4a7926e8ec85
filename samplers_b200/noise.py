"""Noise models with the reference's interface (samplers/noise.py:13-138):
``log_prob(residual)``, ``score(residual)``, ``sample(shape, ...)``, buffers
``sigma`` / ``rate``, properties ``device`` / ``dtype``; ValueError for
non-positive or non-scalar parameters.

Inside the fused DPS step only one number per noise model is used: the
*likelihood weight* w with  d(sum log p)/d(residual) = -w * residual
(``_likelihood_weight``; Gaussian w = 1/sigma^2 from noise.py:78-79, Poisson
approximation w = 2/(rate + 1e-3) from noise.py:123).  It is read once on the
host (the reference's sigma buffer usually lives on the CPU anyway, SURVEY App.
B-11).  ``log_prob`` / ``sample`` remain torch expressions: they are off the hot
path (``sample`` only simulates observations).
"""
from __future__ import annotations

from abc import ABC, abstractmethod

import torch
from torch import nn

from .dtypes import RNG, Device, DType, Shape, Tensor
from .utils.tensor import validate_tensor_is_scalar


class NoiseModel(nn.Module, ABC):
    @abstractmethod
    def log_prob(self, residual: Tensor) -> Tensor: ...

    def score(self, residual: Tensor) -> Tensor:
        """grad of log p wrt the residual (closed form: -w * residual)."""
        return residual * (-self._likelihood_weight())

    @abstractmethod
    def sample(self, shape: Shape, *, device: Device | None = None, dtype: DType = None,
               generator: RNG = None) -> Tensor: ...

    @abstractmethod
    def _likelihood_weight(self) -> float: ...

    def _draw_affine(self, shape: Shape, device, generator: RNG = None) -> tuple[Tensor, float, float]:
        """(raw draw, scale, shift) with ``sample() == scale * raw + shift``: lets psx_observe / psx_add_noise
        fuse the affine map into the observation pass.  Noise models without this decomposition raise."""
        raise NotImplementedError(f"{type(self).__name__} has no (draw, scale, shift) form")

    @property
    def device(self) -> torch.device:
        return next(self.buffers()).device

    @property
    def dtype(self) -> torch.dtype:
        return next(self.buffers()).dtype


def _scalar_buffer(value, name: str, device: Device, dtype: DType) -> Tensor:
    if isinstance(value, Tensor):
        validate_tensor_is_scalar(value, name)
        return value.detach().clone()
    return torch.tensor(float(value), device=device, dtype=dtype or torch.float32)


class GaussianNoise(NoiseModel):
    """eps ~ N(0, sigma^2) i.i.d."""

    sigma: Tensor

    def __init__(self, sigma: float | Tensor, *, device: Device = None, dtype: DType = None) -> None:
        super().__init__()
        t = _scalar_buffer(sigma, "sigma", device, dtype)
        if bool(t <= 0):
            raise ValueError("σ must be positive.")
        self.register_buffer("sigma", t)

    def log_prob(self, r: Tensor) -> Tensor:
        return -(r.square().sum(dim=tuple(range(1, r.ndim)))) / (2 * self.sigma.pow(2))

    def _likelihood_weight(self) -> float:
        var = self.sigma.detach().to(device="cpu", dtype=torch.float32).pow(2)
        return 1.0 / float(var)

    def _draw_affine(self, shape, device, generator=None):
        raw = torch.randn(tuple(shape), dtype=torch.float32, device=device, generator=generator)
        return raw, float(self.sigma.detach().to(device="cpu", dtype=torch.float32)), 0.0

    def sample(self, shape: Shape, *, device: Device = None, dtype: DType = None, generator: RNG = None) -> Tensor:
        device = self.sigma.device if device is None else device
        dtype = self.sigma.dtype if dtype is None else dtype
        return torch.randn(shape, dtype=dtype, device=device, generator=generator) * self.sigma.to(dtype)


class PoissonNoise(NoiseModel):
    """eps = k - rate with k ~ Poisson(rate); Gaussian-style likelihood with variance rate."""

    rate: Tensor

    def __init__(self, rate: float | Tensor, *, device: Device | None = None, dtype: DType = None) -> None:
        super().__init__()
        t = _scalar_buffer(rate, "rate", device, dtype)
        if bool(t <= 0):
            raise ValueError("λ (rate) must be positive.")
        self.register_buffer("rate", t)

    def log_prob(self, r: Tensor) -> Tensor:
        return -(r.pow(2) / (self.rate + 1e-3)).sum(dim=tuple(range(1, r.ndim)))

    def _likelihood_weight(self) -> float:
        lam = self.rate.detach().to(device="cpu", dtype=torch.float32)
        return 2.0 / float(lam + 1e-3)

    def _draw_affine(self, shape, device, generator=None):
        lam = float(self.rate.detach().to(device="cpu", dtype=torch.float32))
        k = torch.poisson(torch.full(tuple(shape), lam, device=device, dtype=torch.float32), generator=generator)
        return k, 1.0, -lam

    def sample(self, shape: Shape, *, device: Device = None, dtype: DType = None, generator: RNG = None) -> Tensor:
        device = self.rate.device if device is None else device
        dtype = self.rate.dtype if dtype is None else dtype
        lam = torch.full(tuple(shape), float(self.rate), device=device, dtype=dtype)
        return torch.poisson(lam, generator=generator) - lam
