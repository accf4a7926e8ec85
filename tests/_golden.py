"""Helpers shared by the golden-vector tests (loads tests/golden/*.npz)."""
from __future__ import annotations

import glob
import json
import os

import numpy as np
import torch

from oracle import operators as oops
from oracle.tiny_net import TinyEpsNet

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_names():
    return sorted(os.path.basename(p)[4:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, "dps_*.npz")))


class Golden:
    def __init__(self, name: str):
        z = np.load(os.path.join(GOLDEN_DIR, f"dps_{name}.npz"))
        self.meta = json.loads(bytes(z["meta"]).decode())
        self.a = {k: torch.from_numpy(z[k]) for k in z.files if k != "meta"}
        self.name = name

    def __getitem__(self, k):
        return self.a[k]

    @property
    def K(self):
        return len(self.meta["t"])

    @property
    def L(self):
        return self.meta["L"]

    @property
    def shape(self):
        return tuple(self.meta["shape"])

    def net(self, device="cpu"):
        net = TinyEpsNet(channels=self.shape[0])
        net.load_state_dict({k[4:]: v for k, v in self.a.items() if k.startswith("net.")})
        return net.to(device)

    def y_flat(self):
        """Observation shaped to broadcast against (L, *y_shape)."""
        y = self.a["y"]
        return y if len(self.meta["batch"]) else y.unsqueeze(0)

    def oracle_op(self):
        spec = self.meta["op"]
        kind = spec[0]
        if kind == "identity":
            return oops.OracleIdentity(self.shape)
        if kind == "mask":
            return oops.OracleMaskGather(self.shape, self.a["mask"])
        if kind == "gblur":
            return oops.OracleSeparableBlur(self.shape, self.a["taps"])
        if kind == "motion":
            return oops.OracleConv2dBlur(self.shape, self.a["kernel2d"])
        if kind == "box":
            return oops.OracleBoxDownsample(self.shape, spec[1])
        raise ValueError(kind)


def rel_err(a: torch.Tensor, b: torch.Tensor) -> float:
    """Per-step state relative error: |a-b|_2 / |b|_2 (the 1e-5 metric)."""
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))


def make_network(g: "Golden", device):
    """Product-side EpsilonNetwork around the golden TinyEpsNet with the fixture's schedule."""
    from samplers_b200.networks.base import EpsilonNetwork

    class GoldenNetwork(EpsilonNetwork):
        def __init__(self, core, acp, ts):
            super().__init__(alphas_cumprod=acp)
            self.core = core
            self._ts = ts

        def forward(self, x, t):
            return self.core(x, int(t))

        @classmethod
        def from_pretrained(cls, *a, **k):
            raise NotImplementedError

        def set_sampling_parameters(self, num_sampling_steps, batch_size=1, num_reconstructions=1):
            self._batch_size, self._num_sampling_steps = batch_size, num_sampling_steps
            self._num_reconstructions = num_reconstructions
            self.register_buffer("timesteps", self._ts.to(self.alphas_cumprod.device))

        @property
        def is_condition_initialized(self):
            return True

    return GoldenNetwork(g.net(), g["acp"].clone(), g["timesteps"].clone()).to(device)


def make_problem(g: "Golden", device):
    """Product-side InverseProblem equivalent to the fixture's."""
    from samplers_b200 import operators as pops
    from samplers_b200.inverse_problem import InverseProblem
    from samplers_b200.noise import GaussianNoise, PoissonNoise

    spec, shape = g.meta["op"], g.shape
    kind = spec[0]
    if kind == "identity":
        op = pops.IdentityOperator(shape)
    elif kind == "mask":
        op = pops.InpaintingOperator(shape, g["mask"])
    elif kind == "gblur":
        op = pops.SeparableBlurOperator(shape, g["taps"])
    elif kind == "motion":
        op = pops.MotionBlurOperator(shape, kernel=g["kernel2d"])
    elif kind == "box":
        op = pops.BoxDownsampleOperator(shape, spec[1])
    else:
        raise ValueError(kind)
    nk, nparam = g.meta["noise"]
    noise = GaussianNoise(sigma=nparam) if nk == "gaussian" else PoissonNoise(rate=nparam)
    return InverseProblem(operator=op.to(device), observation=g["y"].to(device), noise=noise)


# --------------------------------------------------------------------------- PSLD fixtures
def psld_names():
    return sorted(os.path.basename(p)[5:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, "psld_*.npz")))


class PsldGolden:
    def __init__(self, name: str):
        z = np.load(os.path.join(GOLDEN_DIR, f"psld_{name}.npz"))
        self.meta = json.loads(bytes(z["meta"]).decode())
        self.a = {k: torch.from_numpy(z[k]) for k in z.files if k != "meta"}

    def __getitem__(self, k):
        return self.a[k]

    @property
    def K(self):
        return len(self.meta["t"])

    @property
    def L(self):
        return self.meta["L"]

    @property
    def shape(self):
        return tuple(self.meta["shape"])

    def core(self, device="cpu"):
        from oracle.tiny_latent_net import TinyLatentCore
        core = TinyLatentCore(channels=self.shape[0])
        core.load_state_dict({k[4:]: v for k, v in self.a.items() if k.startswith("net.")})
        return core.to(device)

    def oracle_op(self):
        spec = self.meta["op"]
        if spec[0] == "identity":
            return oops.OracleIdentity(self.shape)
        if spec[0] == "gblur":
            return oops.OracleSeparableBlur(self.shape, self.a["taps"])
        if spec[0] == "box":
            return oops.OracleBoxDownsample(self.shape, spec[1])
        if spec[0] == "mask":
            return oops.OracleMaskGather(self.shape, self.a["mask"])
        raise ValueError(spec)

    def y_flat(self, op):
        """psld.py:111 -- observation tiled over reconstructions, flat."""
        y, R = self.a["y"], self.meta["R"]
        nb = len(self.meta["batch"])
        y = y.reshape(-1, *op.y_shape) if nb else y.unsqueeze(0)
        return y.repeat_interleave(R, dim=0)


def make_latent_network(g: "PsldGolden", device):
    from samplers_b200.networks.base import LatentEpsilonNetwork

    class GoldenLatentNetwork(LatentEpsilonNetwork):
        def __init__(self, core, acp, ts):
            super().__init__(alphas_cumprod=acp)
            self.core, self._ts = core, ts

        def forward(self, x, t):
            return self.core.eps(x, int(t))

        @classmethod
        def from_pretrained(cls, *a, **k):
            raise NotImplementedError

        def set_sampling_parameters(self, num_sampling_steps, batch_size=1, num_reconstructions=1):
            self._batch_size, self._num_sampling_steps = batch_size, num_sampling_steps
            self._num_reconstructions = num_reconstructions
            self.register_buffer("timesteps", self._ts.to(self.alphas_cumprod.device))

        @property
        def is_condition_initialized(self):
            return True

        def get_latent_shape(self, x_shape):
            return (4, x_shape[1] // 2, x_shape[2] // 2)

        def _decode(self, z, *, differentiable=False):
            return self.core.decode(z)

        def _encode(self, x, *, differentiable=False):
            return self.core.encode(x)

    return GoldenLatentNetwork(g.core(), g["acp"].clone(), g["timesteps"].clone()).to(device)


def make_psld_problem(g: "PsldGolden", device):
    from samplers_b200 import operators as pops
    from samplers_b200.inverse_problem import InverseProblem
    from samplers_b200.noise import GaussianNoise
    spec, shape = g.meta["op"], g.shape
    if spec[0] == "identity":
        op = pops.IdentityOperator(shape)
    elif spec[0] == "gblur":
        op = pops.SeparableBlurOperator(shape, g["taps"])
    elif spec[0] == "mask":
        op = pops.InpaintingOperator(shape, g["mask"])
    else:
        op = pops.BoxDownsampleOperator(shape, spec[1])
    return InverseProblem(operator=op.to(device), observation=g["y"].to(device), noise=GaussianNoise(sigma=g.meta["sigma"]))


# --------------------------------------------------------------------------- ReSample fixtures
def resample_names():
    return sorted(os.path.basename(p)[9:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, "resample_*.npz")))


class ResampleGolden(PsldGolden):
    def __init__(self, name: str):
        z = np.load(os.path.join(GOLDEN_DIR, f"resample_{name}.npz"))
        self.meta = json.loads(bytes(z["meta"]).decode())
        self.a = {k: torch.from_numpy(z[k]) for k in z.files if k != "meta"}

    def eps_threshold(self) -> float:
        nk, p = self.meta["noise"]
        return float(torch.tensor(p, dtype=torch.float32)) if nk == "gaussian" else 1e-3


def make_resample_problem(g: "ResampleGolden", device):
    from samplers_b200.noise import GaussianNoise, PoissonNoise
    prob = make_psld_problem(_SigmaShim(g), device)
    nk, p = g.meta["noise"]
    prob.noise = GaussianNoise(sigma=p) if nk == "gaussian" else PoissonNoise(rate=p)
    return prob


class _SigmaShim:
    """make_psld_problem reads meta['sigma']; ReSample fixtures store the noise as [kind, param]."""

    def __init__(self, g):
        self._g = g
        self.meta = dict(g.meta, sigma=0.05)
        self.shape = g.shape

    def __getitem__(self, k):
        return self._g[k]


# --------------------------------------------------------------------------- PGDM fixtures
def pgdm_names():
    return sorted(os.path.basename(p)[5:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, "pgdm_*.npz")))


class PgdmGolden(Golden):
    def __init__(self, name: str):
        z = np.load(os.path.join(GOLDEN_DIR, f"pgdm_{name}.npz"))
        self.meta = json.loads(bytes(z["meta"]).decode())
        self.a = {k: torch.from_numpy(z[k]) for k in z.files if k != "meta"}
        self.name = name

    def oracle_op(self):
        spec = self.meta["op"]
        if spec[0] == "mask":  # dense form: y has the shape of x, zeros at the missing pixels
            keep = (~self.a["mask"]).float()

            class Dense(oops.OracleOperator):
                x_shape = y_shape = self.shape

                def apply(self, x):
                    return x * keep

                adjoint = pinv = apply
            return Dense()
        return super().oracle_op()

    def y_flat(self):
        y, R = self.a["y"], self.meta["R"]
        op = self.oracle_op()
        y = y.reshape(-1, *op.y_shape) if len(self.meta["batch"]) else y.unsqueeze(0)
        return y.repeat_interleave(R, dim=0) if len(self.meta["batch"]) else y


def make_pgdm_problem(g: "PgdmGolden", device):
    from samplers_b200 import operators as pops
    from samplers_b200.inverse_problem import InverseProblem
    from samplers_b200.noise import GaussianNoise
    spec, shape = g.meta["op"], g.shape
    if spec[0] == "identity":
        op = pops.IdentityOperator(shape)
    elif spec[0] == "mask":
        op = pops.InpaintingOperator(shape, g["mask"], flatten=False)
    else:
        op = pops.BoxDownsampleOperator(shape, spec[1])
    return InverseProblem(operator=op.to(device), observation=g["y"].to(device), noise=GaussianNoise(sigma=0.05))
