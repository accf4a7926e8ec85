"""PosteriorSampler base (API of samplers/samplers/base.py:7-25)."""
from abc import ABC

from ..dtypes import Shape, Tensor
from ..networks import EpsilonNetwork


class PosteriorSampler(ABC):
    def __init__(self, network: EpsilonNetwork):
        self._epsilon_network = network

    @staticmethod
    def _flatten_leading(x: Tensor, *, x_shape: Shape) -> tuple[Tensor, Shape]:
        lead = x.shape[: x.ndim - len(x_shape)]
        return x.reshape(-1, *x_shape), lead

    @staticmethod
    def _unflatten_leading(x_flat: Tensor, *, batch_shape: tuple[int, ...]) -> Tensor:
        return x_flat.reshape(*batch_shape, *x_flat.shape[1:])
