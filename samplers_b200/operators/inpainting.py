"""Inpainting / outpainting operators (API of samplers/operators/inpainting.py:8-330).

``mask`` is boolean with True = missing pixel (inpainting.py:27).  With
``flatten=True`` (the reference default and its only working mode)
``apply`` returns the kept pixels as a vector of length m (gather, :141-145) and
``apply_transpose`` scatters them back into zeros (:178-187).  ``flatten=False``
zeroes the missing pixels in place (:106-109), which the reference cannot
construct but we can.

Inside the fused DPS step the operator is used in its *dense* form: a uint8
keep-mask multiply on a pre-scattered observation; |r|^2 and A^T r are identical
because masked entries are exactly 0.
"""
from __future__ import annotations

import torch
from torch import Tensor

from .. import _native
from ..dtypes import Device, Shape
from .base import LinearOperator, _numel, native_linear


class InpaintingOperator(LinearOperator):
    def __init__(self, x_shape: Shape, mask: Tensor, flatten: bool = True, device: Device = None):
        x_shape = tuple(int(s) for s in x_shape)
        self.flatten = bool(flatten)
        target = torch.device(device) if device is not None else mask.device
        mask = mask.to(target)
        if mask.dtype != torch.bool:
            mask = mask.ne(0)
        if tuple(mask.shape) != x_shape:
            raise ValueError(f"Mask shape incompatible with x_shape: {tuple(mask.shape)} vs. {x_shape}.")
        kept = torch.nonzero(~mask.flatten(), as_tuple=False).squeeze(1)
        self._m_dim = int(kept.numel())
        self._n_dim = int(mask.numel())
        super().__init__(x_shape=x_shape, device=target)
        self.register_buffer("mask", mask)
        self.register_buffer("_kept_indices", kept)
        self.register_buffer("_singular_values", torch.ones(self._m_dim, dtype=torch.float32, device=target))

    def _infer_y_shape(self, x_shape, device: Device = None):
        return (self._m_dim,) if self.flatten else tuple(x_shape)

    @property
    def shape(self) -> tuple[int, int]:
        """(m, n): kept pixels, total pixels."""
        return self._m_dim, self._n_dim

    def get_singular_values(self) -> Tensor:
        return self._singular_values

    # -- SVD factors (inpainting.py:133-195): V^T gathers, V scatters, U = I
    def apply_V_transpose(self, x: Tensor) -> Tensor:
        flat, lead = self._flat(x, self.x_shape)
        out = _native.gather(flat, self._idx(flat.device), scatter=False, n=self._n_dim)
        return out.reshape(*lead, self._m_dim)

    def apply_V(self, z_kept: Tensor) -> Tensor:
        flat, lead = self._flat(z_kept, (self._m_dim,))
        out = _native.gather(flat, self._idx(flat.device), scatter=True, n=self._n_dim)
        return out.reshape(*lead, *self.x_shape)

    def apply_U(self, z: Tensor) -> Tensor:
        return z

    def apply_U_transpose(self, y: Tensor) -> Tensor:
        return y

    def apply(self, x: Tensor) -> Tensor:
        if self.flatten:
            return self.apply_V_transpose(x)
        flat, lead = self._flat(x, self.x_shape)
        return native_linear(self._native_cached(flat.device), flat, False).reshape(*lead, *self.x_shape)

    def apply_transpose(self, y: Tensor) -> Tensor:
        if self.flatten:
            return self.apply_V(y)
        flat, lead = self._flat(y, self.x_shape)
        return native_linear(self._native_cached(flat.device), flat, True).reshape(*lead, *self.x_shape)

    apply_pseudo_inverse = apply_transpose  # unit singular values: A^+ = A^T (inpainting.py:130)

    # -- native plumbing
    def _idx(self, device) -> Tensor:
        return self._kept_indices.to(device).contiguous()

    def _pinv_gain(self) -> float:
        return 1.0

    def _native(self, device):
        keep = (~self.mask).flatten().to(device=device, dtype=torch.uint8).contiguous()
        return _native.NativeOp.mask(keep)

    def _dense_observation(self, y: Tensor) -> Tensor:
        if not self.flatten:
            return y.reshape(-1, self._n_dim).contiguous()
        flat = y.reshape(-1, self._m_dim).contiguous()
        return _native.gather(flat, self._idx(flat.device), scatter=True, n=self._n_dim)


def get_mask_inpaint_center(image_shape: Shape, start_pct: float = 0.25, end_pct: float = 0.75,
                            device: Device = None) -> Tensor:
    """True inside the central rectangle [start, end) of H and W (inpainting.py:269-297)."""
    if not (0 <= start_pct < end_pct <= 1):
        raise ValueError("start_pct and end_pct must satisfy 0 <= start_pct < end_pct <= 1")
    h, w = image_shape[-2], image_shape[-1]
    mask = torch.zeros(tuple(image_shape), dtype=torch.bool, device=device)
    mask[..., int(h * start_pct):int(h * end_pct), int(w * start_pct):int(w * end_pct)] = True
    return mask


def get_mask_side_painting(image_shape: Shape, pct: float = 0.50, left: bool = True,
                           device: Device = None) -> Tensor:
    """True on a vertical slice of width int(W * pct) at the left / right (inpainting.py:300-330)."""
    if not (0 < pct <= 1):
        raise ValueError("pct must satisfy 0 < pct <= 1")
    w = image_shape[-1]
    width = int(w * pct)
    mask = torch.zeros(tuple(image_shape), dtype=torch.bool, device=device)
    if left:
        mask[..., :width] = True
    else:
        mask[..., -width:] = True
    return mask


def get_mask_random(image_shape: Shape, missing_fraction: float = 0.7, seed: int = 0,
                    per_channel: bool = True, device: Device = None) -> Tensor:
    """Random mask, True = missing with probability ``missing_fraction`` (BASELINE config 3)."""
    if not (0.0 <= missing_fraction < 1.0):
        raise ValueError("missing_fraction must be in [0, 1)")
    g = torch.Generator().manual_seed(seed)
    shape = tuple(image_shape) if per_channel else tuple(image_shape[-2:])
    m = torch.rand(shape, generator=g) < missing_fraction
    if not per_channel:
        m = m.expand(tuple(image_shape)).clone()
    return m.to(device) if device is not None else m


class CenterInpaintingOperator(InpaintingOperator):
    def __init__(self, x_shape: Shape, paint_fraction: float = 0.5, device: Device = None):
        if not (0.0 <= paint_fraction <= 1.0):
            raise ValueError("paint_fraction must be in [0, 1]")
        lo, hi = (1.0 - paint_fraction) / 2.0, (1.0 + paint_fraction) / 2.0
        super().__init__(x_shape, get_mask_inpaint_center(x_shape, lo, hi, device=device), device=device)


class CenterOutpaintingOperator(InpaintingOperator):
    def __init__(self, x_shape: Shape, keep_fraction: float = 0.5, device: Device = None):
        if not (0.0 <= keep_fraction <= 1.0):
            raise ValueError("keep_fraction must be in [0, 1]")
        lo, hi = (1.0 - keep_fraction) / 2.0, (1.0 + keep_fraction) / 2.0
        super().__init__(x_shape, ~get_mask_inpaint_center(x_shape, lo, hi, device=device), device=device)


class SidePaintingOperator(InpaintingOperator):
    def __init__(self, x_shape: Shape, paint_fraction: float = 0.5, left: bool = True, device: Device = None):
        if not (0.0 <= paint_fraction <= 1.0):
            raise ValueError("paint_fraction must be in [0, 1]")
        super().__init__(x_shape, get_mask_side_painting(x_shape, paint_fraction, left, device=device),
                         device=device)


class RandomInpaintingOperator(InpaintingOperator):
    """Random-mask inpainting (70 % missing by default) -- BASELINE.json config 3."""

    def __init__(self, x_shape: Shape, missing_fraction: float = 0.7, seed: int = 0, flatten: bool = True,
                 device: Device = None):
        super().__init__(x_shape, get_mask_random(x_shape, missing_fraction, seed), flatten=flatten,
                         device=device)
