"""Stand-in for the diffusers DDPMScheduler surface the reference touches
(samplers/networks/diffusers/ddpm.py:14, 55-57): ``alphas_cumprod``,
``set_timesteps(n, device=)``, ``timesteps`` (descending).  diffusers is not in
this image; behaviour restated from its published defaults: linear betas
1e-4..0.02 over 1000 train steps, "leading" spacing arange(n) * (1000 // n)."""
from __future__ import annotations

import torch


class DDPMSchedulerLite:
    def __init__(self, num_train_timesteps: int = 1000, beta_start: float = 1e-4, beta_end: float = 0.02):
        self.num_train_timesteps = int(num_train_timesteps)
        betas = torch.linspace(beta_start, beta_end, self.num_train_timesteps, dtype=torch.float32)
        self.alphas_cumprod = torch.cumprod(1.0 - betas, dim=0)
        self.timesteps = torch.arange(self.num_train_timesteps - 1, -1, -1, dtype=torch.int64)

    def set_timesteps(self, num_inference_steps: int, device=None) -> None:
        if num_inference_steps > self.num_train_timesteps:
            raise ValueError("num_inference_steps cannot exceed num_train_timesteps")
        ratio = self.num_train_timesteps // num_inference_steps
        asc = torch.arange(0, num_inference_steps, dtype=torch.int64) * ratio
        self.timesteps = asc.flip(0).to(device)
