"""Tiny deterministic LATENT eps-network for PSLD goldens / parity tests (oracle; TEST INFRA).
eps-net: two 3x3 convs on the latent; encoder: 2x2 stride-2 conv (3 -> 4 channels) + tanh-free affine;
decoder: 2x2 stride-2 transposed conv (4 -> 3) followed by a smooth nonlinearity so that the VAE
Jacobians are not constant."""
from __future__ import annotations

import torch
from torch import nn


class TinyLatentCore(nn.Module):
    def __init__(self, channels: int = 3, latent: int = 4, hidden: int = 8, seed: int = 4321):
        super().__init__()
        g = torch.Generator().manual_seed(seed)
        self.e1 = nn.Conv2d(latent, hidden, 3, padding=1)
        self.e2 = nn.Conv2d(hidden, latent, 3, padding=1)
        self.enc = nn.Conv2d(channels, latent, 2, stride=2)
        self.dec = nn.ConvTranspose2d(latent, channels, 2, stride=2)
        with torch.no_grad():
            for p in self.parameters():
                p.copy_(torch.randn(p.shape, generator=g) * (0.3 if p.ndim > 1 else 0.05))
        self.requires_grad_(False)

    def eps(self, z, t):
        h = torch.nn.functional.silu(self.e1(z))
        return self.e2(h) * (1.0 + 0.001 * float(int(t))) + 0.1 * z

    def encode(self, x):
        return self.enc(x) * 0.5

    def decode(self, z):
        y = self.dec(z)
        return y + 0.1 * torch.tanh(y)
