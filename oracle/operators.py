"""Oracle degradation operators (CPU torch; TEST INFRASTRUCTURE).

Identity and inpainting restate the reference:
  * identity      -> samplers/operators/identity.py:36-66
  * mask (gather) -> samplers/operators/inpainting.py:49-67 (kept indices, mask
                     True = missing), :141-145 (gather), :178-187 (scatter into
                     zeros) and samplers/operators/linear.py:141-165 (unit
                     singular values in between).

Gaussian blur, motion blur and box super-resolution are NOT in the reference
(samplers/operators/__init__.py:1-24 exports none): PARITY UNPINNED for their
arithmetic.  The conventions chosen here are the specification the CUDA kernels
are held to:
  * blur = depthwise cross-correlation, zero padding "same" (k // 2 each side),
    the same taps for every channel; Gaussian taps exp(-k^2 / (2 sigma^2)),
    k = -(K//2)..K//2, normalised to sum 1 in fp64 then rounded to fp32, applied
    separably: horizontal pass, then vertical pass.
  * adjoint of a zero-padded correlation = correlation with the flipped taps.
  * box super-resolution = mean over non-overlapping f x f blocks; adjoint =
    replicate / f^2.
Every operator here passes <A x, y> == <x, A^T y> (tests/test_oracle_operators.py).
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F
from torch import Tensor


class OracleOperator:
    x_shape: tuple
    y_shape: tuple

    def apply(self, x: Tensor) -> Tensor:  # (L, *x_shape) -> (L, *y_shape)
        raise NotImplementedError

    def adjoint(self, y: Tensor) -> Tensor:  # (L, *y_shape) -> (L, *x_shape)
        raise NotImplementedError

    def pinv(self, y: Tensor) -> Tensor:  # Moore-Penrose pseudo-inverse, where it has a closed form
        raise NotImplementedError


class OracleIdentity(OracleOperator):
    def __init__(self, x_shape, flatten: bool = False):
        self.x_shape = tuple(x_shape)
        self.flatten = flatten
        self.y_shape = (int(np.prod(self.x_shape)),) if flatten else self.x_shape

    def apply(self, x):
        return x.reshape(x.shape[0], *self.y_shape)

    def adjoint(self, y):
        return y.reshape(y.shape[0], *self.x_shape)

    pinv = adjoint


class OracleMaskGather(OracleOperator):
    """Keep the pixels where ``mask`` is False, as a flat vector of length m."""

    def __init__(self, x_shape, mask: Tensor):
        self.x_shape = tuple(x_shape)
        mask = mask if mask.dtype == torch.bool else mask.ne(0)
        assert tuple(mask.shape) == self.x_shape
        self.mask = mask
        self.kept = torch.nonzero(~mask.flatten(), as_tuple=False).squeeze(1)
        self.n = mask.numel()
        self.y_shape = (int(self.kept.numel()),)

    def apply(self, x):
        flat = x.reshape(x.shape[0], -1)
        return flat.index_select(1, self.kept) * 1.0

    def adjoint(self, y):
        out = torch.zeros(y.shape[0], self.n, dtype=y.dtype, device=y.device)
        out[:, self.kept] = y * 1.0
        return out.reshape(y.shape[0], *self.x_shape)

    pinv = adjoint  # unit singular values (inpainting.py:130)


def gaussian_taps(kernel_size: int = 61, sigma: float = 3.0) -> Tensor:
    half = kernel_size // 2
    k = np.arange(-half, half + 1, dtype=np.float64)
    w = np.exp(-(k * k) / (2.0 * sigma * sigma))
    w = w / w.sum()
    return torch.from_numpy(w.astype(np.float32))


def motion_line_kernel(kernel_size: int = 61, angle_deg: float = 30.0,
                       length: float | None = None) -> Tensor:
    """Anti-aliased straight-line motion PSF (k x k), normalised to sum 1."""
    half = kernel_size // 2
    length = float(kernel_size - 1) if length is None else float(length)
    k2d = np.zeros((kernel_size, kernel_size), dtype=np.float64)
    steps = max(8 * kernel_size, 64)
    th = math.radians(angle_deg)
    for s in np.linspace(-0.5, 0.5, steps):
        fx = half + s * length * math.cos(th)
        fy = half - s * length * math.sin(th)
        x0, y0 = int(math.floor(fx)), int(math.floor(fy))
        ax, ay = fx - x0, fy - y0
        for yy, wy in ((y0, 1 - ay), (y0 + 1, ay)):
            for xx, wx in ((x0, 1 - ax), (x0 + 1, ax)):
                if 0 <= yy < kernel_size and 0 <= xx < kernel_size:
                    k2d[yy, xx] += wy * wx
    k2d /= k2d.sum()
    return torch.from_numpy(k2d.astype(np.float32))


def motion_walk_kernel(kernel_size: int = 61, intensity: float = 0.5,
                       seed: int = 0) -> Tensor:
    """Random-trajectory motion PSF (camera-shake style), normalised to sum 1.

    A 2-D random walk whose heading changes by N(0, (intensity*pi/2)^2) per step,
    rescaled to fit the k x k window and splatted bilinearly.
    """
    rng = np.random.default_rng(seed)
    n = 4 * kernel_size
    heading = rng.uniform(0, 2 * math.pi)
    pts = np.zeros((n, 2))
    for i in range(1, n):
        heading += rng.normal(0.0, intensity * math.pi / 2 / math.sqrt(kernel_size))
        pts[i] = pts[i - 1] + (math.cos(heading), math.sin(heading))
    pts -= (pts.max(0) + pts.min(0)) / 2
    span = np.abs(pts).max()
    half = kernel_size // 2
    pts = pts / max(span, 1e-9) * (half - 1) * min(1.0, 0.25 + intensity) + half
    k2d = np.zeros((kernel_size, kernel_size))
    for fx, fy in pts:
        x0, y0 = int(math.floor(fx)), int(math.floor(fy))
        ax, ay = fx - x0, fy - y0
        for yy, wy in ((y0, 1 - ay), (y0 + 1, ay)):
            for xx, wx in ((x0, 1 - ax), (x0 + 1, ax)):
                if 0 <= yy < kernel_size and 0 <= xx < kernel_size:
                    k2d[yy, xx] += wy * wx
    k2d /= k2d.sum()
    return torch.from_numpy(k2d.astype(np.float32))


class OracleSeparableBlur(OracleOperator):
    """y = V(H(x)); H correlates rows with ``taps_h``, V columns with ``taps_v``."""

    def __init__(self, x_shape, taps_h: Tensor, taps_v: Tensor | None = None):
        self.x_shape = tuple(x_shape)
        self.y_shape = self.x_shape
        self.taps_h = taps_h.to(torch.float32)
        self.taps_v = self.taps_h if taps_v is None else taps_v.to(torch.float32)
        assert self.taps_h.numel() % 2 == 1 and self.taps_v.numel() % 2 == 1

    @staticmethod
    def _pass(x, taps, horizontal):
        c = x.shape[1]
        k = taps.numel()
        if horizontal:
            w = taps.to(x).view(1, 1, 1, k).expand(c, 1, 1, k)
            return F.conv2d(x, w, padding=(0, k // 2), groups=c)
        w = taps.to(x).view(1, 1, k, 1).expand(c, 1, k, 1)
        return F.conv2d(x, w, padding=(k // 2, 0), groups=c)

    def apply(self, x):
        return self._pass(self._pass(x, self.taps_h, True), self.taps_v, False)

    def adjoint(self, y):
        t = self._pass(y, self.taps_v.flip(0), False)
        return self._pass(t, self.taps_h.flip(0), True)


class OracleGaussianBlur(OracleSeparableBlur):
    def __init__(self, x_shape, kernel_size: int = 61, sigma: float = 3.0):
        super().__init__(x_shape, gaussian_taps(kernel_size, sigma))


class OracleConv2dBlur(OracleOperator):
    """Depthwise 2-D correlation with an arbitrary (k x k) PSF, zero 'same' pad."""

    def __init__(self, x_shape, kernel2d: Tensor):
        self.x_shape = tuple(x_shape)
        self.y_shape = self.x_shape
        self.kernel2d = kernel2d.to(torch.float32)
        assert kernel2d.shape[0] % 2 == 1 and kernel2d.shape[1] % 2 == 1

    def _corr(self, x, k2d):
        c = x.shape[1]
        kh, kw = k2d.shape
        w = k2d.to(x).view(1, 1, kh, kw).expand(c, 1, kh, kw)
        return F.conv2d(x, w, padding=(kh // 2, kw // 2), groups=c)

    def apply(self, x):
        return self._corr(x, self.kernel2d)

    def adjoint(self, y):
        return self._corr(y, self.kernel2d.flip(0, 1))


class OracleBoxDownsample(OracleOperator):
    def __init__(self, x_shape, factor: int = 4):
        self.x_shape = tuple(x_shape)
        c, h, w = self.x_shape
        assert h % factor == 0 and w % factor == 0
        self.factor = factor
        self.y_shape = (c, h // factor, w // factor)

    def apply(self, x):
        return F.avg_pool2d(x, self.factor)

    def adjoint(self, y):
        f = self.factor
        up = y.repeat_interleave(f, dim=-2).repeat_interleave(f, dim=-1)
        return up / float(f * f)

    def pinv(self, y):  # A A^T = I / f^2  =>  A^+ = f^2 A^T = replicate
        f = self.factor
        return y.repeat_interleave(f, dim=-2).repeat_interleave(f, dim=-1)


class OracleMaskedBox(OracleBoxDownsample):
    """BASELINE config 3 read as one composed operator: y = keep * box_f(x), ``keep`` a 0/1 array over the coarse
    grid (dense form, zeros at the dropped pixels).  Neither factor exists in the reference; this
    class is the definition the CUDA path is checked against.  A A^T = diag(keep) / f^2."""

    def __init__(self, x_shape, factor: int, keep: Tensor):
        super().__init__(x_shape, factor)
        assert tuple(keep.shape) == self.y_shape
        self.kept = keep.to(torch.float32)

    def apply(self, x):
        return super().apply(x) * self.kept.to(x)

    def adjoint(self, y):
        return super().adjoint(y * self.kept.to(y))

    def pinv(self, y):
        return super().pinv(y * self.kept.to(y))
