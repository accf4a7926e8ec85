"""Image <-> tensor conversion with the reference's names and arithmetic (samplers/utils/image.py:9-64).

CUDA tensors are converted ON THE DEVICE (psx_image_to_u8 / psx_image_from_u8): the [-1, 1] -> uint8 HWC map of
``tensor_to_pil`` runs before the device->host copy, so a posterior sample crosses PCIe as 1 byte per value instead
of 4, and ``pil_to_tensor(..., device="cuda")`` uploads bytes and expands them on the device.  The arithmetic is the
reference's, rounding for rounding (clamp, +1, *0.5, then torchvision's ``mul(255).byte()``; ``/255``, ``*2``, ``-1``).
"""
from __future__ import annotations

import numpy as np
import torch
from PIL import Image

from .. import _native
from ..dtypes import Device, DType, Tensor


def tensor_to_uint8(img: Tensor) -> Tensor:
    """(..., C, H, W) float32 CUDA tensor in [-1, 1] -> (..., H, W, C) uint8 CUDA tensor."""
    if not img.is_cuda or img.dtype != torch.float32 or img.ndim < 3:
        raise ValueError("tensor_to_uint8 expects a float32 CUDA tensor of shape (..., C, H, W)")
    c, h, w = img.shape[-3:]
    src = img.contiguous()
    images = src.numel() // (c * h * w)
    out = torch.empty((*img.shape[:-3], h, w, c), dtype=torch.uint8, device=img.device)
    _native.image_to_u8(src, out, images, c, h, w)
    return out


def uint8_to_tensor(img: Tensor, dtype: DType = torch.float32) -> Tensor:
    """(..., H, W, C) uint8 CUDA tensor -> (..., C, H, W) tensor in [-1, 1]."""
    if not img.is_cuda or img.dtype != torch.uint8 or img.ndim < 3:
        raise ValueError("uint8_to_tensor expects a uint8 CUDA tensor of shape (..., H, W, C)")
    h, w, c = img.shape[-3:]
    src = img.contiguous()
    images = src.numel() // (c * h * w)
    out = torch.empty((*img.shape[:-3], c, h, w), dtype=torch.float32, device=img.device)
    _native.image_from_u8(src, out, images, c, h, w)
    return out.to(dtype)


def tensor_to_pil(img_tensor: Tensor) -> Image.Image:
    """(C, H, W) or (1, C, H, W) tensor in [-1, 1] -> PIL image (RGB for C = 3, L for C = 1)."""
    if img_tensor.is_cuda:
        u8 = tensor_to_uint8(img_tensor.to(torch.float32)).cpu()
    else:  # host tensors: nothing to offload, same arithmetic in torch
        t = ((img_tensor.clamp(-1.0, 1.0) + 1.0) * 0.5).to(torch.float32)
        u8 = t.mul(255).byte().movedim(-3, -1).contiguous()
    while u8.ndim > 3 and u8.shape[0] == 1:
        u8 = u8[0]
    if u8.ndim != 3:
        raise ValueError("tensor_to_pil expects a single image")
    arr = u8.numpy()
    return Image.fromarray(arr[..., 0], mode="L") if arr.shape[-1] == 1 else Image.fromarray(arr)


def pil_to_tensor(image: Image.Image, device: Device = None, dtype: DType = torch.float32) -> Tensor:
    """PIL image -> (C, H, W) tensor in [-1, 1] on ``device`` (default CPU)."""
    arr = np.asarray(image)
    if arr.ndim == 2:
        arr = arr[..., None]
    if arr.dtype != np.uint8:
        raise ValueError("pil_to_tensor supports 8-bit images")
    u8 = torch.from_numpy(np.array(arr, copy=True))
    device = torch.device(device) if device is not None else torch.device("cpu")
    if device.type == "cuda":
        return uint8_to_tensor(u8.to(device), dtype)
    t = u8.movedim(-1, -3).contiguous().to(torch.float32).div(255)
    return ((t * 2.0) - 1.0).to(dtype)
