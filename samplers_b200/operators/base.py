"""Operator ABCs with the reference's public surface (samplers/operators/base.py:8-113,
samplers/operators/linear.py:11-47): ``x_shape``, ``y_shape``, ``apply``,
``apply_transpose``, ``apply_pseudo_inverse``, ``forward``.

What differs from the reference: the arithmetic of the concrete operators runs in
libpsx (sm_100a kernels) through ``NativeOp``; ``y_shape`` is computed from the
shape instead of by pushing a dummy tensor through ``apply`` (base.py:36-49), so
construction needs no device.  Operators that want the fused DPS path implement
``_native(device)`` and ``_dense_observation(y)``; anything else makes
``DPSSampler`` raise ``NotImplementedError`` (there is no generic / CPU path).
"""
from __future__ import annotations

from abc import ABC, abstractmethod

import torch
from torch import Tensor

from .. import _native
from ..dtypes import Device, Shape


class Operator(torch.nn.Module, ABC):
    """Forward model A of an inverse problem (mandatory ``apply``)."""

    def __init__(self, x_shape: Shape, device: Device = None) -> None:
        super().__init__()
        self.x_shape = tuple(int(s) for s in x_shape)
        self.y_shape = tuple(self._infer_y_shape(self.x_shape, device=device))
        self._native_cache: dict = {}

    @abstractmethod
    def _infer_y_shape(self, x_shape: Shape, device: Device = None) -> Shape: ...

    @abstractmethod
    def apply(self, x: Tensor) -> Tensor:
        """y = A(x); x is (*batch, *x_shape), y is (*batch, *y_shape)."""

    def apply_transpose(self, y: Tensor) -> Tensor:
        raise NotImplementedError("Transpose not defined for this operator")

    def apply_pseudo_inverse(self, y: Tensor) -> Tensor:
        raise NotImplementedError("Pseudo-inverse not defined for this operator")

    def forward(self, x: Tensor) -> Tensor:
        return self.apply(x)

    # ------------------------------------------------------------ native plumbing
    def _native(self, device: torch.device) -> "_native.NativeOp":
        """The libpsx descriptor of this operator on ``device`` (built once per device)."""
        raise NotImplementedError(
            f"{type(self).__name__} has no sm_100a kernel: the fused DPS path supports Identity, "
            "Inpainting, BoxDownsample, GaussianBlur / SeparableBlur and MotionBlur operators only")

    def _native_cached(self, device) -> "_native.NativeOp":
        device = torch.device(device)
        if device.type != "cuda":
            raise RuntimeError(
                f"{type(self).__name__} runs on CUDA tensors only: its arithmetic exists only as sm_100a "
                "kernels (libpsx); there is no CPU path")
        key = (device.type, device.index if device.index is not None else torch.cuda.current_device())
        op = self._native_cache.get(key)
        if op is None:
            with torch.cuda.device(device):
                op = self._native(device)
            self._native_cache[key] = op
        return op

    def _pinv_gain(self) -> float:
        """c with A^+ = c A^T and A A^T = I / c (identity, inpainting: 1; f x f box: f^2).  Operators
        without such a pseudo-inverse cannot be used by PGDM."""
        raise NotImplementedError("Pseudo-inverse not defined for this operator")

    def _dense_observation(self, y: Tensor) -> Tensor:
        """Observation as the kernels index it: (num_obs, n_y) contiguous fp32."""
        return y.reshape(-1, _numel(self.y_shape)).contiguous()

    def _flat(self, t: Tensor, tail: tuple) -> tuple[Tensor, tuple]:
        lead = tuple(t.shape[: t.ndim - len(tail)])
        if tuple(t.shape[t.ndim - len(tail):]) != tuple(tail):
            raise ValueError(f"expected trailing shape {tuple(tail)}, got {tuple(t.shape)}")
        return t.reshape(-1, _numel(tail)).contiguous(), lead


class NonlinearOperator(Operator):
    """Non-linear degradation operator (API parity with operators/base.py:95-113)."""


class LinearOperator(Operator):
    """Linear operator with optional adjoint / pseudo-inverse (linear.py:11-47)."""


def _numel(shape) -> int:
    n = 1
    for s in shape:
        n *= int(s)
    return n


class _NativeLinearFn(torch.autograd.Function):
    """A (or A^T) as an autograd node whose backward is the other one -- lets the
    operator kernels sit inside torch graphs (PSLD / ReSample go through VAEs)."""

    @staticmethod
    def forward(ctx, inp: Tensor, op: "_native.NativeOp", transpose: bool):
        ctx.op, ctx.transpose = op, transpose
        return op.adjoint(inp) if transpose else op.apply(inp)

    @staticmethod
    def backward(ctx, g: Tensor):
        g = g.contiguous()
        out = ctx.op.apply(g) if ctx.transpose else ctx.op.adjoint(g)
        return out, None, None


def native_linear(op: "_native.NativeOp", inp: Tensor, transpose: bool) -> Tensor:
    if inp.requires_grad and torch.is_grad_enabled():
        return _NativeLinearFn.apply(inp, op, transpose)
    return op.adjoint(inp) if transpose else op.apply(inp)
