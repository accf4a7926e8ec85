"""Type aliases used in signatures (same names as the reference's samplers/dtypes.py:7-22)."""
from typing import Sequence, Union

import torch
from torch import Tensor  # noqa: F401

Shape = Union[Sequence[int], torch.Size]
Device = Union[torch.device, str, None]
DType = Union[torch.dtype, None]
Scalars = Union[int, float]
TensorLike = Union[Tensor, int, float]
RNG = Union[torch.Generator, None]
