"""IdentityOperator (API of samplers/operators/identity.py:8-68): A = I, with an
optional flattening of the observation.  Pure views, no kernel needed for
apply / transpose; the fused DPS step uses the PSX_OP_IDENTITY kernel."""
from __future__ import annotations

import torch
from torch import Tensor

from .. import _native
from ..dtypes import Device, Shape
from .base import LinearOperator, _numel


class IdentityOperator(LinearOperator):
    def __init__(self, x_shape: Shape, flatten: bool = False) -> None:
        self.flatten = bool(flatten)
        super().__init__(x_shape=x_shape)

    def _infer_y_shape(self, x_shape, device: Device = None):
        return (_numel(x_shape),) if self.flatten else tuple(x_shape)

    def apply(self, x: Tensor) -> Tensor:
        if not self.flatten:
            return x
        lead = x.shape[: x.ndim - len(self.x_shape)]
        return x.reshape(*lead, *self.y_shape)

    def apply_transpose(self, y: Tensor) -> Tensor:
        if not self.flatten:
            return y
        lead = y.shape[: y.ndim - len(self.y_shape)]
        return y.reshape(*lead, *self.x_shape)

    apply_pseudo_inverse = apply_transpose

    def _pinv_gain(self) -> float:
        return 1.0

    def _native(self, device):
        return _native.NativeOp.identity(_numel(self.x_shape))
