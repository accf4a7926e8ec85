"""DDPMNetwork -- unconditional pixel-space eps-network adapter with the
reference's interface (samplers/networks/diffusers/ddpm.py:12-87): wraps a
DDPM *pipeline* (an object with ``.unet(sample=, timestep=).sample``,
``.scheduler.alphas_cumprod / set_timesteps / timesteps``, ``.to``, ``.device``,
``.dtype``), pads alphas_cumprod with a leading 1.0, exposes ascending
``timesteps``.

``from_pretrained`` needs the third-party ``diffusers`` package + network
access; neither exists in this image, so ``from_config`` builds the same named
architecture with random-init weights behind an ``OfflineDDPMPipeline``.
"""
from __future__ import annotations

from typing import Any

import torch
from torch import Tensor

from ..dtypes import Device, DType
from .base import EpsilonNetwork, NoCondition
from .schedulers import DDPMSchedulerLite
from .unet2d import CELEBAHQ_256, TINY, UNet2DModel

_CONFIGS = {"google/ddpm-celebahq-256": CELEBAHQ_256, "ddpm-celebahq-256": CELEBAHQ_256, "tiny": TINY}


class OfflineDDPMPipeline:
    """Duck-typed stand-in for diffusers.DDPMPipeline (unet + scheduler only)."""

    def __init__(self, unet: torch.nn.Module, scheduler: DDPMSchedulerLite):
        self.unet, self.scheduler = unet, scheduler

    def to(self, device=None, dtype=None):
        self.unet = self.unet.to(device=device, dtype=dtype)
        return self

    @property
    def device(self) -> torch.device:
        return next(self.unet.parameters()).device

    @property
    def dtype(self) -> torch.dtype:
        return next(self.unet.parameters()).dtype


class DDPMNetwork(EpsilonNetwork[NoCondition]):
    def __init__(self, pipeline):
        acp = pipeline.scheduler.alphas_cumprod
        super().__init__(alphas_cumprod=torch.cat([acp.new_tensor([1.0]), acp]))
        self._conditioning: NoCondition | None = None
        self._pipeline = pipeline
        self._pipeline.unet.eval().requires_grad_(False)
        self.alphas_cumprod = self.alphas_cumprod.to(pipeline.device)

    @classmethod
    def from_pretrained(cls, pretrained_model_name_or_path: str, cache_dir: str | None = None,
                        torch_dtype: DType = None, device: Device = None, **pipeline_kwargs: Any) -> "DDPMNetwork":
        try:
            from diffusers import DDPMPipeline  # type: ignore
        except ImportError as e:  # this image: no diffusers, no network
            raise ImportError(
                "DDPMNetwork.from_pretrained needs the `diffusers` package and the checkpoint; "
                "use DDPMNetwork.from_config(...) for a random-init network of the same architecture") from e
        pipe = DDPMPipeline.from_pretrained(pretrained_model_name_or_path, cache_dir=cache_dir,
                                            torch_dtype=torch_dtype, **pipeline_kwargs)
        return cls(pipe.to(device))

    @classmethod
    def from_config(cls, name: str = "google/ddpm-celebahq-256", *, seed: int = 1234,
                    torch_dtype: DType = None, device: Device = None, channels_last: bool = False,
                    **overrides) -> "DDPMNetwork":
        """Random-init network of a named architecture (weights from torch.manual_seed(seed))."""
        if name not in _CONFIGS:
            raise ValueError(f"unknown config {name!r}; known: {sorted(_CONFIGS)}")
        cfg = {**_CONFIGS[name], **overrides}
        rng_state = torch.get_rng_state()
        torch.manual_seed(seed)
        unet = UNet2DModel(**cfg)
        torch.set_rng_state(rng_state)
        pipe = OfflineDDPMPipeline(unet, DDPMSchedulerLite()).to(device=device, dtype=torch_dtype)
        if channels_last:
            pipe.unet = pipe.unet.to(memory_format=torch.channels_last)
        return cls(pipe)

    def forward(self, sample: Tensor, t: Tensor | int) -> Tensor:  # noqa: N802
        if self._num_sampling_steps is None:
            raise RuntimeError("Call `set_sampling_parameters()` before sampling.")
        return self.unet(sample=sample, timestep=t).sample

    def set_sampling_parameters(self, num_sampling_steps: int, batch_size: int = 1, num_reconstructions: int = 1):
        self._batch_size = batch_size
        self._num_sampling_steps = num_sampling_steps
        self._num_reconstructions = num_reconstructions
        self._pipeline.scheduler.set_timesteps(num_sampling_steps, device=self.device)
        # schedulers hand back descending steps; the bridge update wants s < t < ell ascending
        self.register_buffer("timesteps", torch.flip(self._pipeline.scheduler.timesteps, dims=(0,)), persistent=True)

    def is_condition_initialized(self) -> bool:  # method, as in the reference
        return True

    @property
    def unet(self):
        return self._pipeline.unet

    def to(self, device: torch.device | str | None = None, dtype: torch.dtype | None = None):
        device = torch.device(device) if device is not None else self.device
        dtype = dtype if dtype is not None else self.dtype
        super().to(device=device, dtype=dtype)
        self._pipeline = self._pipeline.to(device=device, dtype=dtype)
        return self

    @property
    def device(self) -> torch.device:
        return self._pipeline.device

    @property
    def dtype(self) -> torch.dtype:
        return self._pipeline.dtype
