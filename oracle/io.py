"""Oracle for the data formats either side of the sampling loop (CPU torch; TEST INFRASTRUCTURE -- only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline may import this).

Restates, op for op (each torch op is one fp32 rounding):
  * image_to_u8       samplers/utils/image.py:24-32  clamp(-1,1); (x+1)*0.5; then torchvision
                      transforms.functional.to_pil_image (0.26): ``pic.mul(255).byte()`` and CHW -> HWC
  * image_from_u8     samplers/utils/image.py:57-64  torchvision to_tensor: HWC uint8 -> CHW float32 ``.div(255)``;
                      then ``tensor * 2.0 - 1.0``
  * simulate_observation  samplers/inverse_problem.py:55-62 (``y_clean + eps``) with
                      GaussianNoise.sample noise.py:91-92 (``randn * sigma``) or
                      PoissonNoise.sample noise.py:134-138 (``poisson(lam) - lam``); the raw draw is an input.
Pinned against the unmodified reference by oracle/make_golden_io.py -> tests/golden/io_*.npz.
"""
from __future__ import annotations

import torch
from torch import Tensor


def image_to_u8(x: Tensor) -> Tensor:
    """(..., C, H, W) float in [-1, 1] -> (..., H, W, C) uint8."""
    t = (x.clamp(-1.0, 1.0) + 1.0) * 0.5
    return t.to(torch.float32).mul(255).byte().movedim(-3, -1).contiguous()


def image_from_u8(u8: Tensor) -> Tensor:
    """(..., H, W, C) uint8 -> (..., C, H, W) float32 in [-1, 1]."""
    t = u8.movedim(-1, -3).contiguous().to(torch.float32).div(255)
    return (t * 2.0) - 1.0


def simulate_observation(clean: Tensor, noise_kind: str, param: float, raw: Tensor) -> Tensor:
    """y = clean + eps; raw is the N(0,1) draw (gaussian, param = sigma) or the Poisson count k (poisson, param = rate)."""
    if noise_kind == "gaussian":
        eps = raw * torch.tensor(param, dtype=clean.dtype)
    elif noise_kind == "poisson":
        eps = raw - torch.full(raw.shape, float(param), dtype=clean.dtype)
    else:
        raise ValueError(noise_kind)
    return clean + eps
