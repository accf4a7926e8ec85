"""Noise schedule + timestep grid restatement (oracle; test infrastructure).

The reference obtains both from the third-party ``diffusers`` package, which is
absent from /root/reference and from this image (unpinned in
requirements.txt:2).  Restated from its published behaviour:

* ``DDPMScheduler(num_train_timesteps=1000, beta_start=1e-4, beta_end=0.02,
  beta_schedule="linear")``: betas = linspace in fp32, alphas_cumprod =
  cumprod(1 - betas) in fp32.
* ``set_timesteps(N)`` with the default ``timestep_spacing="leading"``:
  ``arange(N) * (1000 // N)`` reversed (descending), int64.

Reference call sites: samplers/networks/diffusers/ddpm.py:13-20 (pads the
table with a leading 1.0), samplers/networks/base.py:23-26 (clips to
[1e-6, 1]), samplers/networks/diffusers/ddpm.py:45-58 (flips to ascending).
"""
from __future__ import annotations

import torch


def ddpm_linear_alphas_cumprod(num_train_timesteps: int = 1000,
                               beta_start: float = 1e-4,
                               beta_end: float = 0.02) -> torch.Tensor:
    betas = torch.linspace(beta_start, beta_end, num_train_timesteps, dtype=torch.float32)
    return torch.cumprod(1.0 - betas, dim=0)


def padded_clipped_acp(acp: torch.Tensor) -> torch.Tensor:
    """ddpm.py:14-17 then networks/base.py:25."""
    one = acp.new_tensor([1.0])
    return torch.cat([one, acp]).clip(1e-6, 1)


def leading_timesteps_ascending(num_sampling_steps: int,
                                num_train_timesteps: int = 1000) -> torch.Tensor:
    """Ascending grid as registered by ddpm.py:55-58."""
    ratio = num_train_timesteps // num_sampling_steps
    return torch.arange(0, num_sampling_steps, dtype=torch.int64) * ratio
