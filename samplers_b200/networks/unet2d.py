"""Pure-torch UNet with the architecture of diffusers' ``UNet2DModel`` as
configured for ``google/ddpm-celebahq-256``: 3 -> 3 channels,
block_out_channels (128, 128, 256, 256, 512, 512), 2 layers per block, single-head
self-attention in the 5th down block / 2nd up block and the mid block, GroupNorm(32,
eps 1e-6), SiLU, sinusoidal time embedding (flip_sin_to_cos=False, freq_shift=1),
downsample_padding 0.  ~113.7 M parameters.

There is no network access for the checkpoint, so weights are random-init; this
module only has to be the *named architecture* for throughput realism.  It stays
in torch (cuDNN convolutions, cuBLAS / SDPA attention): the north star keeps the
eps prediction and its VJP behind the network abstraction.
"""
from __future__ import annotations

import dataclasses
import math

import torch
import torch.nn.functional as F
from torch import Tensor, nn

CELEBAHQ_256 = dict(
    in_channels=3, out_channels=3, block_out_channels=(128, 128, 256, 256, 512, 512),
    layers_per_block=2, attn_down=(False, False, False, False, True, False),
    attn_up=(False, True, False, False, False, False), norm_num_groups=32, norm_eps=1e-6,
    sample_size=256,
)
# A narrow variant with the same topology for tests / smoke runs.
TINY = dict(
    in_channels=3, out_channels=3, block_out_channels=(32, 32, 64), layers_per_block=1,
    attn_down=(False, True, False), attn_up=(False, True, False), norm_num_groups=8, norm_eps=1e-6,
    sample_size=32,
)


@dataclasses.dataclass
class UNet2DOutput:
    sample: Tensor


def timestep_embedding(t: Tensor, dim: int, freq_shift: float = 1.0) -> Tensor:
    half = dim // 2
    exponent = -math.log(10000.0) * torch.arange(half, dtype=torch.float32, device=t.device) / (half - freq_shift)
    arg = t.float()[:, None] * torch.exp(exponent)[None, :]
    return torch.cat([torch.sin(arg), torch.cos(arg)], dim=-1)


class ResnetBlock(nn.Module):
    def __init__(self, cin: int, cout: int, temb: int, groups: int, eps: float):
        super().__init__()
        self.norm1 = nn.GroupNorm(groups, cin, eps=eps)
        self.conv1 = nn.Conv2d(cin, cout, 3, padding=1)
        self.time_emb_proj = nn.Linear(temb, cout)
        self.norm2 = nn.GroupNorm(groups, cout, eps=eps)
        self.conv2 = nn.Conv2d(cout, cout, 3, padding=1)
        self.conv_shortcut = nn.Conv2d(cin, cout, 1) if cin != cout else None

    def forward(self, x: Tensor, temb: Tensor) -> Tensor:
        h = self.conv1(F.silu(self.norm1(x)))
        h = h + self.time_emb_proj(F.silu(temb))[:, :, None, None]
        h = self.conv2(F.silu(self.norm2(h)))
        return (x if self.conv_shortcut is None else self.conv_shortcut(x)) + h


class SelfAttention(nn.Module):
    def __init__(self, ch: int, groups: int, eps: float):
        super().__init__()
        self.group_norm = nn.GroupNorm(groups, ch, eps=eps)
        self.to_q, self.to_k, self.to_v = nn.Linear(ch, ch), nn.Linear(ch, ch), nn.Linear(ch, ch)
        self.to_out = nn.Linear(ch, ch)

    def forward(self, x: Tensor) -> Tensor:
        b, c, h, w = x.shape
        t = self.group_norm(x).reshape(b, c, h * w).transpose(1, 2)
        q, k, v = self.to_q(t)[:, None], self.to_k(t)[:, None], self.to_v(t)[:, None]  # one head
        o = F.scaled_dot_product_attention(q, k, v)[:, 0]
        return x + self.to_out(o).transpose(1, 2).reshape(b, c, h, w)


class DownBlock(nn.Module):
    def __init__(self, cin, cout, temb, layers, attn, add_down, groups, eps):
        super().__init__()
        self.resnets = nn.ModuleList(ResnetBlock(cin if i == 0 else cout, cout, temb, groups, eps)
                                     for i in range(layers))
        self.attentions = nn.ModuleList(SelfAttention(cout, groups, eps) for _ in range(layers)) if attn else None
        self.down = nn.Conv2d(cout, cout, 3, stride=2, padding=0) if add_down else None

    def forward(self, x, temb, skips):
        for i, res in enumerate(self.resnets):
            x = res(x, temb)
            if self.attentions is not None:
                x = self.attentions[i](x)
            skips.append(x)
        if self.down is not None:
            x = self.down(F.pad(x, (0, 1, 0, 1)))  # downsample_padding = 0
            skips.append(x)
        return x


class UpBlock(nn.Module):
    def __init__(self, cin, cprev, cout, temb, layers, attn, add_up, groups, eps):
        super().__init__()
        res = []
        for i in range(layers):
            skip = cin if i == layers - 1 else cout
            res.append(ResnetBlock((cprev if i == 0 else cout) + skip, cout, temb, groups, eps))
        self.resnets = nn.ModuleList(res)
        self.attentions = nn.ModuleList(SelfAttention(cout, groups, eps) for _ in range(layers)) if attn else None
        self.up = nn.Conv2d(cout, cout, 3, padding=1) if add_up else None

    def forward(self, x, temb, skips):
        for i, res in enumerate(self.resnets):
            x = res(torch.cat([x, skips.pop()], dim=1), temb)
            if self.attentions is not None:
                x = self.attentions[i](x)
        if self.up is not None:
            x = self.up(F.interpolate(x, scale_factor=2.0, mode="nearest"))
        return x


class UNet2DModel(nn.Module):
    def __init__(self, in_channels=3, out_channels=3, block_out_channels=(128, 128, 256, 256, 512, 512),
                 layers_per_block=2, attn_down=None, attn_up=None, norm_num_groups=32, norm_eps=1e-6,
                 sample_size=256):
        super().__init__()
        ch = tuple(block_out_channels)
        nb = len(ch)
        attn_down = tuple(attn_down) if attn_down is not None else (False,) * nb
        attn_up = tuple(attn_up) if attn_up is not None else (False,) * nb
        self.sample_size = sample_size
        self.in_channels = in_channels
        temb = ch[0] * 4
        self.time_dim = ch[0]
        self.conv_in = nn.Conv2d(in_channels, ch[0], 3, padding=1)
        self.time_embedding = nn.Sequential(nn.Linear(ch[0], temb), nn.SiLU(), nn.Linear(temb, temb))
        downs, cout = [], ch[0]
        for i in range(nb):
            cin, cout = cout, ch[i]
            downs.append(DownBlock(cin, cout, temb, layers_per_block, attn_down[i], i < nb - 1,
                                   norm_num_groups, norm_eps))
        self.down_blocks = nn.ModuleList(downs)
        self.mid_res1 = ResnetBlock(ch[-1], ch[-1], temb, norm_num_groups, norm_eps)
        self.mid_attn = SelfAttention(ch[-1], norm_num_groups, norm_eps)
        self.mid_res2 = ResnetBlock(ch[-1], ch[-1], temb, norm_num_groups, norm_eps)
        rev = ch[::-1]
        ups, cout = [], rev[0]
        for i in range(nb):
            cprev, cout = cout, rev[i]
            cin = rev[min(i + 1, nb - 1)]
            ups.append(UpBlock(cin, cprev, cout, temb, layers_per_block + 1, attn_up[i], i < nb - 1,
                               norm_num_groups, norm_eps))
        self.up_blocks = nn.ModuleList(ups)
        self.conv_norm_out = nn.GroupNorm(norm_num_groups, ch[0], eps=norm_eps)
        self.conv_out = nn.Conv2d(ch[0], out_channels, 3, padding=1)

    def forward(self, sample: Tensor, timestep) -> UNet2DOutput:
        t = torch.as_tensor(timestep, device=sample.device)
        t = t.reshape(-1).expand(sample.shape[0]) if t.ndim == 0 or t.numel() == 1 else t
        temb = self.time_embedding(timestep_embedding(t, self.time_dim).to(sample.dtype))
        x = self.conv_in(sample)
        skips = [x]
        for blk in self.down_blocks:
            x = blk(x, temb, skips)
        x = self.mid_res2(self.mid_attn(self.mid_res1(x, temb)), temb)
        for blk in self.up_blocks:
            x = blk(x, temb, skips)
        return UNet2DOutput(sample=self.conv_out(F.silu(self.conv_norm_out(x))))
