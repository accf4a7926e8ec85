"""Box-downsample super-resolution operator (BASELINE.json config 3; absent from
the reference).  A = mean over non-overlapping f x f blocks, A^T = replicate / f^2,
A^+ = replicate (A A^T = I / f^2).  Convention: oracle/operators.py."""
from __future__ import annotations

import torch
from torch import Tensor

from .. import _native
from ..dtypes import Device, Shape
from .base import LinearOperator, native_linear


class BoxDownsampleOperator(LinearOperator):
    def __init__(self, x_shape: Shape, factor: int = 4, device: Device = None):
        x_shape = tuple(int(s) for s in x_shape)
        if len(x_shape) != 3:
            raise ValueError(f"x_shape must be (C, H, W), got {x_shape}")
        if factor < 1 or x_shape[1] % factor or x_shape[2] % factor:
            raise ValueError("H and W must be multiples of factor")
        self.factor = int(factor)
        super().__init__(x_shape=x_shape, device=device)

    def _infer_y_shape(self, x_shape, device: Device = None):
        c, h, w = x_shape
        return (c, h // self.factor, w // self.factor)

    def apply(self, x: Tensor) -> Tensor:
        flat, lead = self._flat(x, self.x_shape)
        return native_linear(self._native_cached(flat.device), flat, False).reshape(*lead, *self.y_shape)

    def apply_transpose(self, y: Tensor) -> Tensor:
        flat, lead = self._flat(y, self.y_shape)
        return native_linear(self._native_cached(flat.device), flat, True).reshape(*lead, *self.x_shape)

    def apply_pseudo_inverse(self, y: Tensor) -> Tensor:
        return self.apply_transpose(y) * float(self.factor * self.factor)

    def _pinv_gain(self) -> float:
        return float(self.factor * self.factor)

    def _native(self, device):
        c, h, w = self.x_shape
        return _native.NativeOp.box(c, h, w, self.factor)


SuperResolutionOperator = BoxDownsampleOperator


class MaskedBoxDownsampleOperator(BoxDownsampleOperator):
    """BASELINE config 3 read as ONE operator: y = keep * box_f(x) -- the ``factor`` x ``factor`` box average followed
    by a pixel mask on the coarse grid (dense form: zeros where ``mask`` is True = missing, the convention of
    ``InpaintingOperator``).  The reference has neither factor; the arithmetic is defined by
    ``oracle/operators.py: OracleMaskedBox``.  K1 is the box kernel with the residual zeroed at dropped pixels."""

    def __init__(self, x_shape: Shape, factor: int = 4, mask: Tensor | None = None, missing_fraction: float = 0.7,
                 seed: int = 0, device: Device = None):
        super().__init__(x_shape, factor, device=device)
        target = torch.device(device) if device is not None else (mask.device if mask is not None else torch.device("cpu"))
        if mask is None:
            g = torch.Generator().manual_seed(int(seed))
            mask = torch.rand(self.y_shape, generator=g) < float(missing_fraction)
        mask = mask.to(target)
        if mask.dtype != torch.bool:
            mask = mask.ne(0)
        if tuple(mask.shape) != tuple(self.y_shape):
            raise ValueError(f"Mask shape incompatible with the coarse grid: {tuple(mask.shape)} vs. {self.y_shape}.")
        self.register_buffer("mask", mask)

    def _native(self, device):
        c, h, w = self.x_shape
        keep = (~self.mask).to(device=device, dtype=torch.uint8).contiguous()
        return _native.NativeOp.box_masked(c, h, w, self.factor, keep)
