"""Blur operators -- named by BASELINE.json's north_star but absent from the
reference (samplers/operators/__init__.py:1-24); they follow the reference's
``LinearOperator`` contract (operators/linear.py:11-47) and the conventions of
oracle/operators.py: depthwise zero-padded "same" cross-correlation, identical
taps for every channel.

* ``SeparableBlurOperator``  y = V(H(x))  -- rows then columns (psx_op_create_sepblur)
* ``GaussianBlurOperator``   separable with taps exp(-k^2 / 2 sigma^2) / sum (61 taps, sigma 3 by default)
* ``MotionBlurOperator``     arbitrary k x k PSF, evaluated row segment by row segment (psx_op_create_conv2d)
"""
from __future__ import annotations

import math

import numpy as np
import torch
from torch import Tensor

from .. import _native
from ..dtypes import Device, Shape
from .base import LinearOperator, native_linear


def gaussian_taps(kernel_size: int = 61, sigma: float = 3.0) -> Tensor:
    if kernel_size < 1 or kernel_size % 2 == 0:
        raise ValueError("kernel_size must be a positive odd integer")
    if sigma <= 0:
        raise ValueError("sigma must be positive")
    k = np.arange(-(kernel_size // 2), kernel_size // 2 + 1, dtype=np.float64)
    w = np.exp(-(k * k) / (2.0 * float(sigma) ** 2))
    return torch.from_numpy((w / w.sum()).astype(np.float32))


def motion_line_kernel(kernel_size: int = 61, angle_deg: float = 30.0, length: float | None = None) -> Tensor:
    """Anti-aliased straight-line PSF through the window centre, normalised to sum 1."""
    if kernel_size < 1 or kernel_size % 2 == 0:
        raise ValueError("kernel_size must be a positive odd integer")
    half = kernel_size // 2
    length = float(kernel_size - 1) if length is None else float(length)
    psf = np.zeros((kernel_size, kernel_size), dtype=np.float64)
    th = math.radians(angle_deg)
    for s in np.linspace(-0.5, 0.5, max(8 * kernel_size, 64)):
        _splat(psf, half + s * length * math.cos(th), half - s * length * math.sin(th))
    return torch.from_numpy((psf / psf.sum()).astype(np.float32))


def motion_walk_kernel(kernel_size: int = 61, intensity: float = 0.5, seed: int = 0) -> Tensor:
    """Camera-shake style PSF: a random walk whose heading jitters with ``intensity``."""
    if kernel_size < 1 or kernel_size % 2 == 0:
        raise ValueError("kernel_size must be a positive odd integer")
    rng = np.random.default_rng(seed)
    n = 4 * kernel_size
    heading = rng.uniform(0, 2 * math.pi)
    pts = np.zeros((n, 2))
    for i in range(1, n):
        heading += rng.normal(0.0, intensity * math.pi / 2 / math.sqrt(kernel_size))
        pts[i] = pts[i - 1] + (math.cos(heading), math.sin(heading))
    pts -= (pts.max(0) + pts.min(0)) / 2
    half = kernel_size // 2
    pts = pts / max(np.abs(pts).max(), 1e-9) * (half - 1) * min(1.0, 0.25 + intensity) + half
    psf = np.zeros((kernel_size, kernel_size))
    for fx, fy in pts:
        _splat(psf, fx, fy)
    return torch.from_numpy((psf / psf.sum()).astype(np.float32))


def _splat(psf: np.ndarray, fx: float, fy: float) -> None:
    k = psf.shape[0]
    x0, y0 = int(math.floor(fx)), int(math.floor(fy))
    ax, ay = fx - x0, fy - y0
    for yy, wy in ((y0, 1 - ay), (y0 + 1, ay)):
        for xx, wx in ((x0, 1 - ax), (x0 + 1, ax)):
            if 0 <= yy < k and 0 <= xx < k:
                psf[yy, xx] += wy * wx


class _ImageOperator(LinearOperator):
    """Shared plumbing for operators on (C, H, W) images with y_shape == x_shape."""

    def __init__(self, x_shape: Shape, device: Device = None):
        if len(tuple(x_shape)) != 3:
            raise ValueError(f"x_shape must be (C, H, W), got {tuple(x_shape)}")
        super().__init__(x_shape=x_shape, device=device)

    def _infer_y_shape(self, x_shape, device: Device = None):
        return tuple(x_shape)

    def apply(self, x: Tensor) -> Tensor:
        flat, lead = self._flat(x, self.x_shape)
        return native_linear(self._native_cached(flat.device), flat, False).reshape(*lead, *self.y_shape)

    def apply_transpose(self, y: Tensor) -> Tensor:
        flat, lead = self._flat(y, self.y_shape)
        return native_linear(self._native_cached(flat.device), flat, True).reshape(*lead, *self.x_shape)


class SeparableBlurOperator(_ImageOperator):
    def __init__(self, x_shape: Shape, taps_h: Tensor, taps_v: Tensor | None = None, device: Device = None):
        taps_v = taps_h if taps_v is None else taps_v
        for name, t in (("taps_h", taps_h), ("taps_v", taps_v)):
            if t.ndim != 1 or t.numel() % 2 == 0:
                raise ValueError(f"{name} must be a 1-D tensor with an odd number of taps")
            if t.numel() > 127:
                raise ValueError(f"{name}: at most 127 taps are supported")
        super().__init__(x_shape, device=device)
        self.register_buffer("taps_h", taps_h.detach().to(torch.float32).cpu().clone())
        self.register_buffer("taps_v", taps_v.detach().to(torch.float32).cpu().clone())

    def _native(self, device):
        c, h, w = self.x_shape
        return _native.NativeOp.sepblur(c, h, w, self.taps_h.tolist(), self.taps_v.tolist())


class GaussianBlurOperator(SeparableBlurOperator):
    """61 x 61, sigma = 3.0 by default (BASELINE.json config 2)."""

    def __init__(self, x_shape: Shape, kernel_size: int = 61, sigma: float = 3.0, device: Device = None):
        self.kernel_size, self.sigma = int(kernel_size), float(sigma)
        super().__init__(x_shape, gaussian_taps(kernel_size, sigma), device=device)


class MotionBlurOperator(_ImageOperator):
    """Depthwise 2-D PSF.  Pass ``kernel`` (k x k) or let it draw a straight line
    (``angle_deg``) / a random camera-shake trajectory (``intensity``, ``seed``)."""

    def __init__(self, x_shape: Shape, kernel: Tensor | None = None, kernel_size: int = 61,
                 angle_deg: float | None = None, intensity: float = 0.5, seed: int = 0,
                 device: Device = None):
        if kernel is None:
            kernel = (motion_line_kernel(kernel_size, angle_deg) if angle_deg is not None
                      else motion_walk_kernel(kernel_size, intensity, seed))
        if kernel.ndim != 2 or kernel.shape[0] % 2 == 0 or kernel.shape[1] % 2 == 0:
            raise ValueError("kernel must be 2-D with odd sizes")
        if max(kernel.shape) > 127:
            raise ValueError("kernel sizes above 127 are not supported")
        super().__init__(x_shape, device=device)
        self.register_buffer("kernel", kernel.detach().to(torch.float32).cpu().clone())

    def _native(self, device):
        c, h, w = self.x_shape
        return _native.NativeOp.conv2d(c, h, w, self.kernel)
