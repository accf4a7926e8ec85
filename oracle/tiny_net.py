"""Tiny deterministic eps-network used by golden vectors and parity tests
(oracle; TEST INFRASTRUCTURE).  Two 3x3 convolutions with a SiLU and a
timestep-dependent gain; weights come from an explicit state dict so the same
function can be rebuilt on any device."""
from __future__ import annotations

import torch
from torch import nn


class TinyEpsNet(nn.Module):
    def __init__(self, channels: int = 3, hidden: int = 8, seed: int = 1234):
        super().__init__()
        g = torch.Generator().manual_seed(seed)
        self.c1 = nn.Conv2d(channels, hidden, 3, padding=1)
        self.c2 = nn.Conv2d(hidden, channels, 3, padding=1)
        with torch.no_grad():
            for p in self.parameters():
                p.copy_(torch.randn(p.shape, generator=g) * (0.3 if p.ndim > 1 else 0.05))
        self.requires_grad_(False)

    def forward(self, x, t):
        gain = 1.0 + 0.001 * float(int(t))
        h = torch.nn.functional.silu(self.c1(x))
        return self.c2(h) * gain + 0.1 * x
