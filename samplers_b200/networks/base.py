"""EpsilonNetwork / LatentEpsilonNetwork -- the torch boundary of the sampler,
with the reference's contract (samplers/networks/base.py:13-117):

* ``alphas_cumprod`` buffer, padded (index 0 -> 1.0, index k -> alpha_bar_k) and
  clipped to [1e-6, 1];
* ``timesteps`` buffer ascending; ``t`` passed to forward / predict_x0 is a buffer index;
* ``forward`` (the eps prediction) is whatever torch module the subclass wraps --
  it stays in torch (cuDNN / cuBLAS), by design of the north star;
* ``predict_x0`` is the Tweedie estimate; on CUDA it runs psx_tweedie.
"""
from __future__ import annotations

import dataclasses
from abc import ABC, abstractmethod
from typing import Generic, TypeVar

import torch
from torch import Tensor

from .. import _native
from ..dtypes import Shape

C = TypeVar("C")


def tweedie_scalars(alphas_cumprod: Tensor, t: int) -> tuple[float, float]:
    """(sqrt(acp_t), sqrt(1 - acp_t)) with torch's fp32 roundings (base.py:42-43)."""
    a = alphas_cumprod[int(t)].detach().to(device="cpu", dtype=torch.float32)
    return float(a ** 0.5), float((1 - a) ** 0.5)


class _TweedieFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x: Tensor, eps: Tensor, sa: float, s1: float):
        ctx.sa, ctx.s1 = sa, s1
        xc, ec = x.contiguous(), eps.contiguous()
        _native.require_cuda(xc, "x")
        _native.require_cuda(ec, "eps")
        out = torch.empty_like(xc)
        L = xc.shape[0] if xc.ndim > 1 else 1
        _native.tweedie(xc.view(L, -1), ec.view(L, -1), sa, s1, out.view(L, -1))
        return out

    @staticmethod
    def backward(ctx, g: Tensor):
        gx = g / ctx.sa
        return gx, gx * (-ctx.s1), None, None


class EpsilonNetwork(torch.nn.Module, ABC, Generic[C]):
    def __init__(self, alphas_cumprod: Tensor):
        super().__init__()
        self.register_buffer("alphas_cumprod", alphas_cumprod.clip(1e-6, 1))
        self._batch_size = None
        self._num_sampling_steps = None
        self._num_reconstructions = None

    @abstractmethod
    def forward(self, x: Tensor, t: Tensor | int): ...

    def predict_noise(self, x: Tensor, t: Tensor | int):
        return self.forward(x, t)

    def predict_x0(self, x: Tensor, t: Tensor | int):
        sa, s1 = tweedie_scalars(self.alphas_cumprod, int(t))
        return _TweedieFn.apply(x, self.forward(x, t), sa, s1)

    def score(self, x: Tensor, t: Tensor):
        acp_t = self.alphas_cumprod[t]
        return -self.forward(x, t) / ((1 - acp_t) ** 0.5)

    @property
    def device(self) -> torch.device:
        return self.alphas_cumprod.device

    @property
    def dtype(self) -> torch.dtype:
        return self.alphas_cumprod.dtype

    @classmethod
    @abstractmethod
    def from_pretrained(cls, *args, **kwargs): ...

    @abstractmethod
    def set_sampling_parameters(self, num_sampling_steps: int, batch_size: int = 1,
                                num_reconstructions: int = 1): ...

    @property
    def are_sampling_parameters_initialized(self) -> bool:
        return self._batch_size is not None

    def clear_sampling_parameters(self):
        self._batch_size = None
        self._num_sampling_steps = None
        self._num_reconstructions = None

    def set_condition(self, condition: C | None) -> None: ...

    @property
    @abstractmethod
    def is_condition_initialized(self) -> bool: ...

    def clear_condition(self): ...


class LatentEpsilonNetwork(EpsilonNetwork[C], ABC, Generic[C]):
    @abstractmethod
    def get_latent_shape(self, x_shape: Shape) -> Shape: ...

    def decode(self, z: Tensor, differentiable: bool = False):
        if differentiable:
            return self._decode(z=z, differentiable=True)
        with torch.no_grad():
            return self._decode(z=z, differentiable=False).detach()

    @abstractmethod
    def _decode(self, z: Tensor, *, differentiable: bool = False): ...

    def encode(self, x: Tensor, differentiable: bool = False):
        if differentiable:
            return self._encode(x=x, differentiable=True)
        with torch.no_grad():
            return self._encode(x=x, differentiable=False).detach()

    @abstractmethod
    def _encode(self, x: Tensor, *, differentiable: bool = False): ...


@dataclasses.dataclass(slots=True)
class NoCondition:
    """Marker: this diffusion prior takes no conditioning."""
