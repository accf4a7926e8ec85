"""StableDiffusionNetwork -- latent eps-network adapter with the reference's interface
(samplers/networks/diffusers/stable_diffusion.py:49-362): wraps a Stable-Diffusion *pipeline* (``.unet``, ``.vae``,
``.scheduler``), pads alphas_cumprod with a leading 1.0, exposes ascending ``timesteps``, caches the conditioning
(``set_condition``), applies classifier-free guidance in ``forward`` (:296-330), and converts between pixel and latent
space with the VAE and its scaling factor (``_decode`` :332-338, ``_encode`` :340-347, latent = distribution *mean*).

``from_pretrained`` needs ``diffusers`` + checkpoints + the CLIP text encoder; none exist in this image.  ``from_config``
builds the SD-1.5 *architecture* in pure torch with random-init weights: UNet2DCondition with
block_out_channels (320, 640, 1280, 1280), 2 layers per block, cross-attention (dim 768, 8 heads... head_dim = C / 8)
in the first three down / last three up blocks and the mid block, GEGLU feed-forward; AutoencoderKL with
(128, 256, 512, 512), 2 layers per block, mid-block attention, 4 latent channels, scaling factor 0.18215, 8x
spatial reduction; scaled-linear betas 0.00085..0.012.  Without a text encoder, prompts are given as embeddings
(``StableDiffusionCondition.prompt_embeds``, (B, 77, 768)); ``None`` means a zero ("null") embedding.  The modules
stay in torch (cuDNN / cuBLAS / SDPA): the north star keeps eps, its VJP and the VAE behind the network abstraction --
this file exists so that PSLD / ReSample can be run at BASELINE.json's configs 4-5 with a realistic cost structure.
"""
from __future__ import annotations

import dataclasses
from typing import Any

import torch
import torch.nn.functional as F
from torch import Tensor, nn

from ..dtypes import Device, DType, Shape
from .base import LatentEpsilonNetwork
from .unet2d import ResnetBlock, timestep_embedding

SD15 = dict(
    unet=dict(in_channels=4, out_channels=4, block_out_channels=(320, 640, 1280, 1280), layers_per_block=2,
              cross_attention_dim=768, num_heads=8, cross_attn=(True, True, True, False), norm_num_groups=32),
    vae=dict(in_channels=3, latent_channels=4, block_out_channels=(128, 256, 512, 512), layers_per_block=2,
             norm_num_groups=32, scaling_factor=0.18215),
    context_len=77,
)
# Same topology, narrow: for tests and smoke runs.
SD15_TINY = dict(
    unet=dict(in_channels=4, out_channels=4, block_out_channels=(32, 64), layers_per_block=1,
              cross_attention_dim=32, num_heads=4, cross_attn=(True, False), norm_num_groups=8),
    vae=dict(in_channels=3, latent_channels=4, block_out_channels=(16, 32, 32, 32), layers_per_block=1,
             norm_num_groups=8, scaling_factor=0.18215),
    context_len=7,
)
_CONFIGS = {"runwayml/stable-diffusion-v1-5": SD15, "sd15": SD15, "sd15-tiny": SD15_TINY}


# ------------------------------------------------------------------------------------------------ transformer
class Attention(nn.Module):
    def __init__(self, dim: int, ctx_dim: int | None, heads: int):
        super().__init__()
        self.heads = heads
        self.to_q = nn.Linear(dim, dim, bias=False)
        self.to_k = nn.Linear(ctx_dim or dim, dim, bias=False)
        self.to_v = nn.Linear(ctx_dim or dim, dim, bias=False)
        self.to_out = nn.Linear(dim, dim)

    def forward(self, x: Tensor, ctx: Tensor | None = None) -> Tensor:
        ctx = x if ctx is None else ctx
        b, n, c = x.shape
        q = self.to_q(x).view(b, n, self.heads, -1).transpose(1, 2)
        k = self.to_k(ctx).view(b, ctx.shape[1], self.heads, -1).transpose(1, 2)
        v = self.to_v(ctx).view(b, ctx.shape[1], self.heads, -1).transpose(1, 2)
        out = F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(b, n, c)
        return self.to_out(out)


class TransformerBlock(nn.Module):
    """LayerNorm -> self-attention, LayerNorm -> cross-attention, LayerNorm -> GEGLU feed-forward (x4)."""

    def __init__(self, dim: int, ctx_dim: int, heads: int):
        super().__init__()
        self.norm1, self.attn1 = nn.LayerNorm(dim), Attention(dim, None, heads)
        self.norm2, self.attn2 = nn.LayerNorm(dim), Attention(dim, ctx_dim, heads)
        self.norm3 = nn.LayerNorm(dim)
        self.ff_in, self.ff_out = nn.Linear(dim, 8 * dim), nn.Linear(4 * dim, dim)

    def forward(self, x: Tensor, ctx: Tensor) -> Tensor:
        x = x + self.attn1(self.norm1(x))
        x = x + self.attn2(self.norm2(x), ctx)
        h, gate = self.ff_in(self.norm3(x)).chunk(2, dim=-1)
        return x + self.ff_out(h * F.gelu(gate))


class SpatialTransformer(nn.Module):
    def __init__(self, ch: int, ctx_dim: int, heads: int, groups: int):
        super().__init__()
        self.norm = nn.GroupNorm(groups, ch, eps=1e-6)
        self.proj_in, self.proj_out = nn.Conv2d(ch, ch, 1), nn.Conv2d(ch, ch, 1)
        self.block = TransformerBlock(ch, ctx_dim, heads)

    def forward(self, x: Tensor, ctx: Tensor) -> Tensor:
        b, c, h, w = x.shape
        y = self.proj_in(self.norm(x)).flatten(2).transpose(1, 2)
        y = self.block(y, ctx).transpose(1, 2).reshape(b, c, h, w)
        return x + self.proj_out(y)


# ------------------------------------------------------------------------------------------------ UNet
@dataclasses.dataclass
class UNetOutput:
    sample: Tensor


class UNet2DConditionLite(nn.Module):
    def __init__(self, in_channels, out_channels, block_out_channels, layers_per_block, cross_attention_dim, num_heads,
                 cross_attn, norm_num_groups, norm_eps: float = 1e-5):
        super().__init__()
        ch, g = tuple(block_out_channels), norm_num_groups
        self.time_dim = ch[0]
        temb = 4 * ch[0]
        self.time_embedding = nn.Sequential(nn.Linear(ch[0], temb), nn.SiLU(), nn.Linear(temb, temb))
        self.conv_in = nn.Conv2d(in_channels, ch[0], 3, padding=1)

        def attn(c, on):
            return SpatialTransformer(c, cross_attention_dim, num_heads, g) if on else None

        self.down_res, self.down_attn, self.down_sample = nn.ModuleList(), nn.ModuleList(), nn.ModuleList()
        skip_ch, cin = [ch[0]], ch[0]
        for i, cout in enumerate(ch):
            for _ in range(layers_per_block):
                self.down_res.append(ResnetBlock(cin, cout, temb, g, norm_eps))
                self.down_attn.append(attn(cout, cross_attn[i]) or nn.Identity())
                cin = cout
                skip_ch.append(cin)
            last = i == len(ch) - 1
            self.down_sample.append(nn.Identity() if last else nn.Conv2d(cin, cin, 3, stride=2, padding=1))
            if not last:
                skip_ch.append(cin)
        self.layers_per_block, self.n_blocks, self.cross_attn = layers_per_block, len(ch), tuple(cross_attn)
        self.mid_res1 = ResnetBlock(cin, cin, temb, g, norm_eps)
        self.mid_attn = SpatialTransformer(cin, cross_attention_dim, num_heads, g)
        self.mid_res2 = ResnetBlock(cin, cin, temb, g, norm_eps)
        self.up_res, self.up_attn, self.up_sample = nn.ModuleList(), nn.ModuleList(), nn.ModuleList()
        for i, cout in enumerate(reversed(ch)):
            on = tuple(reversed(cross_attn))[i]
            for _ in range(layers_per_block + 1):
                self.up_res.append(ResnetBlock(cin + skip_ch.pop(), cout, temb, g, norm_eps))
                self.up_attn.append(attn(cout, on) or nn.Identity())
                cin = cout
            last = i == len(ch) - 1
            self.up_sample.append(nn.Identity() if last else nn.Conv2d(cin, cin, 3, padding=1))
        self.conv_norm_out = nn.GroupNorm(g, ch[0], eps=norm_eps)
        self.conv_out = nn.Conv2d(ch[0], out_channels, 3, padding=1)

    def forward(self, sample: Tensor, timestep, encoder_hidden_states: Tensor, **_unused) -> UNetOutput:
        t = torch.as_tensor(timestep, device=sample.device)
        t = t.reshape(-1).expand(sample.shape[0]) if t.ndim == 0 or t.numel() == 1 else t
        temb = self.time_embedding(timestep_embedding(t, self.time_dim, freq_shift=0.0).to(sample.dtype))
        ctx = encoder_hidden_states.to(sample.dtype)
        x = self.conv_in(sample)
        skips, k = [x], 0
        for i in range(self.n_blocks):
            for _ in range(self.layers_per_block):
                x = self.down_res[k](x, temb)
                x = self.down_attn[k](x, ctx) if self.cross_attn[i] else x
                skips.append(x)
                k += 1
            if i < self.n_blocks - 1:
                x = self.down_sample[i](x)
                skips.append(x)
        x = self.mid_res2(self.mid_attn(self.mid_res1(x, temb), ctx), temb)
        k = 0
        for i in range(self.n_blocks):
            on = self.cross_attn[self.n_blocks - 1 - i]
            for _ in range(self.layers_per_block + 1):
                x = self.up_res[k](torch.cat([x, skips.pop()], dim=1), temb)
                x = self.up_attn[k](x, ctx) if on else x
                k += 1
            if i < self.n_blocks - 1:
                x = self.up_sample[i](F.interpolate(x, scale_factor=2.0, mode="nearest"))
        return UNetOutput(sample=self.conv_out(F.silu(self.conv_norm_out(x))))


# ------------------------------------------------------------------------------------------------ VAE
class _VaeAttention(nn.Module):
    def __init__(self, ch: int, groups: int):
        super().__init__()
        self.norm, self.attn = nn.GroupNorm(groups, ch, eps=1e-6), Attention(ch, None, 1)

    def forward(self, x: Tensor) -> Tensor:
        b, c, h, w = x.shape
        y = self.attn(self.norm(x).flatten(2).transpose(1, 2)).transpose(1, 2).reshape(b, c, h, w)
        return x + y


class _VaeRes(ResnetBlock):
    """ResnetBlock without a time embedding."""

    def __init__(self, cin: int, cout: int, groups: int):
        super().__init__(cin, cout, 1, groups, 1e-6)
        self.time_emb_proj = None

    def forward(self, x: Tensor, temb=None) -> Tensor:
        h = self.conv1(F.silu(self.norm1(x)))
        h = self.conv2(F.silu(self.norm2(h)))
        return (x if self.conv_shortcut is None else self.conv_shortcut(x)) + h


@dataclasses.dataclass
class DiagonalGaussian:
    mean: Tensor
    logvar: Tensor


class AutoencoderKLLite(nn.Module):
    def __init__(self, in_channels, latent_channels, block_out_channels, layers_per_block, norm_num_groups,
                 scaling_factor):
        super().__init__()
        ch, g = tuple(block_out_channels), norm_num_groups
        self.scaling_factor, self.latent_channels = float(scaling_factor), latent_channels
        self.scale = 2 ** (len(ch) - 1)
        enc: list[nn.Module] = [nn.Conv2d(in_channels, ch[0], 3, padding=1)]
        cin = ch[0]
        for i, cout in enumerate(ch):
            for _ in range(layers_per_block):
                enc.append(_VaeRes(cin, cout, g))
                cin = cout
            if i < len(ch) - 1:
                enc.append(nn.Conv2d(cin, cin, 3, stride=2, padding=1))
        enc += [_VaeRes(cin, cin, g), _VaeAttention(cin, g), _VaeRes(cin, cin, g), nn.GroupNorm(g, cin, eps=1e-6),
                nn.SiLU(), nn.Conv2d(cin, 2 * latent_channels, 3, padding=1), nn.Conv2d(2 * latent_channels,
                                                                                          2 * latent_channels, 1)]
        self.encoder = nn.Sequential(*enc)
        dec: list[nn.Module] = [nn.Conv2d(latent_channels, latent_channels, 1), nn.Conv2d(latent_channels, cin, 3, padding=1),
                                _VaeRes(cin, cin, g), _VaeAttention(cin, g), _VaeRes(cin, cin, g)]
        for i, cout in enumerate(reversed(ch)):
            for _ in range(layers_per_block + 1):
                dec.append(_VaeRes(cin, cout, g))
                cin = cout
            if i < len(ch) - 1:
                dec += [nn.Upsample(scale_factor=2.0, mode="nearest"), nn.Conv2d(cin, cin, 3, padding=1)]
        dec += [nn.GroupNorm(g, cin, eps=1e-6), nn.SiLU(), nn.Conv2d(cin, in_channels, 3, padding=1)]
        self.decoder = nn.Sequential(*dec)

    def encode(self, x: Tensor) -> DiagonalGaussian:
        mean, logvar = self.encoder(x).chunk(2, dim=1)
        return DiagonalGaussian(mean=mean, logvar=logvar.clamp(-30.0, 20.0))

    def decode(self, z: Tensor) -> Tensor:
        return self.decoder(z)


# ------------------------------------------------------------------------------------------------ pipeline + adapter
class ScaledLinearScheduler:
    """The scheduler surface the adapter touches; SD-1.5 defaults: scaled-linear betas 0.00085..0.012, 1000 steps,
    leading spacing with steps_offset 1; ``scale_model_input`` is the identity for this family."""

    def __init__(self, num_train_timesteps: int = 1000, beta_start: float = 0.00085, beta_end: float = 0.012,
                 steps_offset: int = 1):
        self.num_train_timesteps, self.steps_offset = int(num_train_timesteps), int(steps_offset)
        betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, self.num_train_timesteps, dtype=torch.float32) ** 2
        self.alphas_cumprod = torch.cumprod(1.0 - betas, dim=0)
        self.timesteps = torch.arange(self.num_train_timesteps - 1, -1, -1, dtype=torch.int64)

    def set_timesteps(self, num_inference_steps: int, device=None) -> None:
        if num_inference_steps > self.num_train_timesteps:
            raise ValueError("num_inference_steps cannot exceed num_train_timesteps")
        ratio = self.num_train_timesteps // num_inference_steps
        asc = torch.arange(0, num_inference_steps, dtype=torch.int64) * ratio + self.steps_offset
        self.timesteps = asc.clamp_(max=self.num_train_timesteps - 1).flip(0).to(device)

    def scale_model_input(self, sample: Tensor, t) -> Tensor:
        return sample


class OfflineSDPipeline:
    """Duck-typed stand-in for diffusers.StableDiffusionPipeline (unet + vae + scheduler; no text encoder)."""

    def __init__(self, unet: nn.Module, vae: AutoencoderKLLite, scheduler, context_len: int, context_dim: int):
        self.unet, self.vae, self.scheduler = unet, vae, scheduler
        self.context_len, self.context_dim = context_len, context_dim
        self.vae_scale_factor = vae.scale

    def to(self, device=None, dtype=None):
        self.unet = self.unet.to(device=device, dtype=dtype)
        self.vae = self.vae.to(device=device, dtype=dtype)
        return self

    @property
    def device(self) -> torch.device:
        return next(self.unet.parameters()).device

    @property
    def dtype(self) -> torch.dtype:
        return next(self.unet.parameters()).dtype


@dataclasses.dataclass(slots=True)
class StableDiffusionCondition:
    """The fields of the reference's condition object that can be honoured without a text encoder.  Defaults as in the
    reference (networks/diffusers/stable_diffusion.py:16-28): classifier-free guidance is ON (7.5) unless the caller
    passes guidance_scale <= 1."""
    prompt: str | list[str] = ""
    negative_prompt: str | list[str] | None = None
    guidance_scale: float = 7.5
    guidance_rescale: float = 0.0
    prompt_embeds: Tensor | None = None            # (B or 1, context_len, context_dim)
    negative_prompt_embeds: Tensor | None = None
    cross_attention_kwargs: dict[str, Any] | None = None


@dataclasses.dataclass(slots=True)
class ConditioningState:
    prompt_embeds: Tensor
    do_classifier_free_guidance: bool
    guidance_scale: float
    guidance_rescale: float


class StableDiffusionNetwork(LatentEpsilonNetwork[StableDiffusionCondition]):
    def __init__(self, pipeline) -> None:
        acp = pipeline.scheduler.alphas_cumprod
        super().__init__(alphas_cumprod=torch.cat([acp.new_tensor([1.0]), acp]))
        self._conditioning: ConditioningState | None = None
        self._pipeline = pipeline
        self._pipeline.unet.eval().requires_grad_(False)
        self._pipeline.vae.eval().requires_grad_(False)
        self._vae_latent_multiplier = self.vae.scaling_factor
        self.latent_resolution_ratio = self._pipeline.vae_scale_factor
        self.latent_num_channels = self.vae.latent_channels
        self.alphas_cumprod = self.alphas_cumprod.to(pipeline.device)

    @classmethod
    def from_pretrained(cls, pretrained_model_name_or_path: str, cache_dir: str | None = None,
                        torch_dtype: DType = None, device: Device = None, **pipeline_kwargs: Any):
        try:
            from diffusers import StableDiffusionPipeline  # type: ignore
        except ImportError as e:  # this image: no diffusers, no network
            raise ImportError(
                "StableDiffusionNetwork.from_pretrained needs the `diffusers` package and the checkpoint; use "
                "StableDiffusionNetwork.from_config(...) for a random-init network of the same architecture") from e
        pipe = StableDiffusionPipeline.from_pretrained(pretrained_model_name_or_path, cache_dir=cache_dir,
                                                       torch_dtype=torch_dtype, **pipeline_kwargs)
        return cls(pipe.to(device))

    @classmethod
    def from_config(cls, name: str = "runwayml/stable-diffusion-v1-5", *, seed: int = 1234, torch_dtype: DType = None,
                    device: Device = None) -> "StableDiffusionNetwork":
        """Random-init network of a named architecture (weights from torch.manual_seed(seed))."""
        if name not in _CONFIGS:
            raise ValueError(f"unknown config {name!r}; known: {sorted(_CONFIGS)}")
        cfg = _CONFIGS[name]
        rng_state = torch.get_rng_state()
        torch.manual_seed(seed)
        unet, vae = UNet2DConditionLite(**cfg["unet"]), AutoencoderKLLite(**cfg["vae"])
        torch.set_rng_state(rng_state)
        pipe = OfflineSDPipeline(unet, vae, ScaledLinearScheduler(), cfg["context_len"],
                                 cfg["unet"]["cross_attention_dim"]).to(device=device, dtype=torch_dtype)
        return cls(pipe)

    # ---------------------------------------------------------------- sampler-facing API
    def set_sampling_parameters(self, num_sampling_steps: int, batch_size: int = 1, num_reconstructions: int = 1):
        self._batch_size = batch_size
        self._num_sampling_steps = num_sampling_steps
        self._num_reconstructions = num_reconstructions
        self._pipeline.scheduler.set_timesteps(num_sampling_steps, device=self.device)
        self.register_buffer("timesteps", torch.flip(self._pipeline.scheduler.timesteps, dims=(0,)), persistent=True)

    def get_latent_shape(self, x_shape: Shape) -> Shape:
        channel, height, width = x_shape
        f = self.latent_resolution_ratio
        if (height % f) or (width % f):
            raise ValueError(f"H={height} and W={width} must both be divisible by the latent_scale_factor={f}.")
        return self.latent_num_channels, height // f, width // f

    @torch.no_grad()
    def set_condition(self, condition: StableDiffusionCondition | None) -> None:
        if not self.are_sampling_parameters_initialized:
            raise RuntimeError("Call `set_sampling_parameters()` before conditioning.")
        condition = condition if condition is not None else StableDiffusionCondition()
        if (isinstance(condition.prompt, str) and condition.prompt) or (isinstance(condition.prompt, list) and any(condition.prompt)):
            if condition.prompt_embeds is None:
                raise NotImplementedError("no text encoder in this build: pass `prompt_embeds` instead of `prompt`")
        pipe = self._pipeline
        shape = (1, pipe.context_len, pipe.context_dim)
        embeds = condition.prompt_embeds
        embeds = torch.zeros(shape) if embeds is None else embeds
        if embeds.ndim != 3 or tuple(embeds.shape[1:]) != shape[1:]:
            raise ValueError(f"prompt_embeds must have shape (B, {shape[1]}, {shape[2]})")
        total = self._batch_size * self._num_reconstructions
        embeds = embeds.to(device=self.device, dtype=self.dtype)
        if embeds.shape[0] not in (1, self._batch_size, total):
            raise ValueError("prompt_embeds batch must be 1, batch_size or batch_size * num_reconstructions")
        if embeds.shape[0] == self._batch_size and total != self._batch_size:
            embeds = embeds.repeat_interleave(self._num_reconstructions, dim=0)
        embeds = embeds.expand(total, -1, -1) if embeds.shape[0] == 1 else embeds
        cfg = condition.guidance_scale > 1.0
        if cfg:
            neg = condition.negative_prompt_embeds
            neg = torch.zeros_like(embeds) if neg is None else neg.to(embeds).expand_as(embeds)
            embeds = torch.cat([neg, embeds])
        self._conditioning = ConditioningState(prompt_embeds=embeds.contiguous(), do_classifier_free_guidance=cfg,
                                               guidance_scale=float(condition.guidance_scale),
                                               guidance_rescale=float(condition.guidance_rescale))

    @property
    def is_condition_initialized(self) -> bool:
        return self._conditioning is not None

    def clear_condition(self):
        self._conditioning = None

    def forward(self, latents: Tensor, t: Tensor | int) -> Tensor:
        """eps(x_t, t) with classifier-free guidance (stable_diffusion.py:296-330)."""
        if not self.is_condition_initialized:
            raise RuntimeError("Call `set_condition()` before sampling.")
        state = self._conditioning
        x = torch.cat([latents] * 2) if state.do_classifier_free_guidance else latents
        x = self._pipeline.scheduler.scale_model_input(x, t)
        noise = self.unet(sample=x, timestep=t, encoder_hidden_states=state.prompt_embeds).sample
        if state.do_classifier_free_guidance:
            uncond, text = noise.chunk(2)
            noise = uncond + state.guidance_scale * (text - uncond)
            if state.guidance_rescale > 0.0:  # Lin et al. 2023, sec. 3.4 (diffusers rescale_noise_cfg)
                dims = tuple(range(1, noise.ndim))
                rescaled = noise * (text.std(dim=dims, keepdim=True) / noise.std(dim=dims, keepdim=True))
                noise = state.guidance_rescale * rescaled + (1 - state.guidance_rescale) * noise
        return noise

    def _decode(self, z: Tensor, *, differentiable: bool = False) -> Tensor:
        return self.vae.decode(z / self._vae_latent_multiplier)

    def _encode(self, x: Tensor, *, differentiable: bool = False) -> Tensor:
        return self.vae.encode(x).mean * self._vae_latent_multiplier

    @property
    def unet(self):
        return self._pipeline.unet

    @property
    def vae(self):
        return self._pipeline.vae

    def to(self, device: torch.device | str | None = None, dtype: torch.dtype | None = None):
        device = torch.device(device) if device is not None else self.device
        dtype = dtype if dtype is not None else self.dtype
        super().to(device=device)           # the schedule buffer stays fp32: only its device follows
        self._pipeline = self._pipeline.to(device=device, dtype=dtype)
        return self

    @property
    def device(self) -> torch.device:
        return self._pipeline.device

    @property
    def dtype(self) -> torch.dtype:
        return self._pipeline.dtype
