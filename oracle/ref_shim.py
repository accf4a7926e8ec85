"""Import the UNMODIFIED reference from /root/reference (oracle; TEST INFRA).

Works only where /root/reference exists (the build container).  Nothing that
runs on the GPU box (``-m gpu`` tests, smoke(), bench.py) may call this; it is
used by ``oracle/make_golden.py`` to record golden vectors and by CPU tests that
are skipped when the reference tree is absent.

The reference eagerly imports the third-party ``diffusers`` package
(samplers/networks/__init__.py:3-4 -> networks/diffusers/ddpm.py:4,
stable_diffusion.py:5-8), which is not installed.  We pre-seed ``sys.modules``
with empty stand-ins exporting the imported names; no reference file is
modified or copied.
"""
from __future__ import annotations

import contextlib
import os
import sys
import types
from typing import Callable

import torch

REFERENCE_ROOT = "/root/reference"
REAL_RANDN = torch.randn  # the unpatched generator, for draw() callbacks


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "samplers"))


def _stub(name: str, **attrs):
    mod = types.ModuleType(name)
    mod.__dict__.update(attrs)
    mod.__path__ = []  # behave like a package
    sys.modules[name] = mod
    return mod


def load_reference():
    """Returns the reference ``samplers`` package (imported once)."""
    if "samplers" in sys.modules and getattr(sys.modules["samplers"], "__file__", "") \
            .startswith(REFERENCE_ROOT):
        return sys.modules["samplers"]
    if not reference_available():
        raise RuntimeError("reference tree not present")
    if "diffusers" not in sys.modules:
        class _Missing:  # names the reference only uses in annotations / from_pretrained
            pass
        _stub("diffusers", AutoencoderKL=_Missing, StableDiffusionPipeline=_Missing,
              UNet2DConditionModel=_Missing)
        _stub("diffusers.image_processor", PipelineImageInput=_Missing)
        _stub("diffusers.models")
        _stub("diffusers.models.autoencoders")
        _stub("diffusers.models.autoencoders.vae", DiagonalGaussianDistribution=_Missing)
        _stub("diffusers.pipelines")
        _stub("diffusers.pipelines.ddpm")
        _stub("diffusers.pipelines.ddpm.pipeline_ddpm", DDPMPipeline=_Missing, UNet2DModel=_Missing)
        _stub("diffusers.pipelines.stable_diffusion")
        _stub("diffusers.pipelines.stable_diffusion.pipeline_stable_diffusion",
              rescale_noise_cfg=lambda *a, **k: None)
    sys.path.insert(0, REFERENCE_ROOT)
    try:
        import samplers  # noqa: F401  (the reference package)
        import samplers.samplers  # noqa: F401
        import samplers.networks  # noqa: F401
        import samplers.operators  # noqa: F401
    finally:
        sys.path.remove(REFERENCE_ROOT)
    return sys.modules["samplers"]


# ----------------------------------------------------------------------------
# duck-typed DDPM pipeline (what DDPMNetwork touches: ddpm.py:13-20, 43, 55-57, 68-87)
# ----------------------------------------------------------------------------
class _Out:
    def __init__(self, sample):
        self.sample = sample


class _UNetAdapter(torch.nn.Module):
    def __init__(self, fn_module: torch.nn.Module):
        super().__init__()
        self.inner = fn_module

    def forward(self, sample, timestep):
        return _Out(self.inner(sample, timestep))


class _Scheduler:
    def __init__(self, acp: torch.Tensor, num_train: int, custom_ascending=None):
        self.alphas_cumprod = acp
        self.num_train = num_train
        self.timesteps = None
        self._custom = custom_ascending

    def set_timesteps(self, n, device=None):
        if self._custom is not None:
            asc = torch.as_tensor(self._custom(n), dtype=torch.int64)
        else:
            asc = torch.arange(0, n, dtype=torch.int64) * (self.num_train // n)
        self.timesteps = asc.flip(0).to(device)


class FakeDDPMPipeline:
    def __init__(self, net_module: torch.nn.Module, acp: torch.Tensor, num_train: int = 1000,
                 custom_ascending: Callable | None = None):
        self.unet = _UNetAdapter(net_module)
        self.scheduler = _Scheduler(acp, num_train, custom_ascending)
        self.device = torch.device("cpu")
        self.dtype = torch.float32

    def to(self, device=None, dtype=None):
        return self


# ----------------------------------------------------------------------------
# noise injection + per-step recording around the unmodified DPSSampler
# ----------------------------------------------------------------------------
@contextlib.contextmanager
def injected_noise(draw: Callable[[tuple], torch.Tensor]):
    """Replace torch.randn / torch.randn_like (dps.py:83, bridge_kernels.py:59)."""
    real_randn, real_like = torch.randn, torch.randn_like

    def randn(*size, **kw):
        if "size" in kw:
            shape = tuple(kw["size"])
        elif len(size) == 1 and not isinstance(size[0], int):
            shape = tuple(size[0])
        else:
            shape = tuple(size)
        return draw(shape).to(dtype=kw.get("dtype") or torch.float32)

    def randn_like(t, **kw):
        return draw(tuple(t.shape)).to(dtype=t.dtype)

    torch.randn, torch.randn_like = randn, randn_like
    try:
        yield
    finally:
        torch.randn, torch.randn_like = real_randn, real_like


class RecordingNet(torch.nn.Module):
    """Wraps the eps-module and keeps every (x_t, t, eps) the sampler asked for."""

    def __init__(self, inner: torch.nn.Module):
        super().__init__()
        self.inner = inner
        self.calls: list[dict] = []

    def forward(self, x, t):
        eps = self.inner(x, t)
        self.calls.append({"x_t": x.detach().clone(), "t": int(t), "eps": eps.detach().clone()})
        return eps


@contextlib.contextmanager
def recorded_autograd(sink: list):
    real = torch.autograd.grad

    def grad(*a, **k):
        out = real(*a, **k)
        sink.append(out[0].detach().clone())
        return out

    torch.autograd.grad = grad
    try:
        yield
    finally:
        torch.autograd.grad = real
