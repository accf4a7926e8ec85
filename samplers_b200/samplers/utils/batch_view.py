"""BatchView: shape bookkeeping between the structured view
(*batch_shape, num_samples, *data_shape) and the flat view (leading_size, *data_shape)
the kernels work on.  Same properties / methods as the reference helper
(samplers/samplers/utils/batch_view.py:8-146); host-side only, no kernel."""
from __future__ import annotations

import math
from typing import Sequence

import torch
from torch import Tensor


def _as_tuple(v) -> tuple[int, ...]:
    return (int(v),) if isinstance(v, int) else tuple(int(s) for s in v)


class BatchView:
    def __init__(self, batch_shape: int | Sequence[int] | torch.Size, num_samples: int,
                 data_shape: int | Sequence[int] | torch.Size) -> None:
        self._batch_shape = _as_tuple(batch_shape)
        self._num_samples = int(num_samples)
        self._data_shape = _as_tuple(data_shape)

    batch_shape = property(lambda self: self._batch_shape)
    num_samples = property(lambda self: self._num_samples)
    data_shape = property(lambda self: self._data_shape)

    @property
    def batch_size(self) -> int:
        return math.prod(self._batch_shape)

    @property
    def leading_shape(self) -> tuple[int, ...]:
        return (*self._batch_shape, self._num_samples)

    @property
    def leading_size(self) -> int:
        return math.prod(self.leading_shape)

    @property
    def flat_shape(self) -> tuple[int, ...]:
        return (self.leading_size, *self._data_shape)

    @property
    def shape(self) -> tuple[int, ...]:
        return (*self.leading_shape, *self._data_shape)

    @property
    def per_sample_broadcast_shape(self) -> tuple[int, ...]:
        return (self.leading_size,) + (1,) * len(self._data_shape)

    def _tail(self, x: Tensor) -> tuple[int, ...]:
        return tuple(x.shape[x.ndim - len(self._data_shape):])

    def flatten(self, x: Tensor) -> Tensor:
        return x.reshape(self.leading_size, *self._tail(x))

    def unflatten(self, x: Tensor) -> Tensor:
        return x.reshape(*self.leading_shape, *self._tail(x))

    def repeat_observation(self, observation: Tensor) -> Tensor:
        """(*batch_shape, *tail) -> (leading_size, *tail): sample l sees observation l // num_samples."""
        tail = self._tail(observation)
        nb = len(self._batch_shape)
        return self.flatten(observation.unsqueeze(nb).expand(*self.leading_shape, *tail))

    def __repr__(self) -> str:
        return (f"{type(self).__name__}(batch_shape={self._batch_shape}, num_samples={self._num_samples}, "
                f"data_shape={self._data_shape})")
