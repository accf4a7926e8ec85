"""Record PSLD golden vectors from the UNMODIFIED reference PSLDSampler (build container only).
Usage: python -m oracle.make_golden_psld      ->  tests/golden/psld_<case>.npz"""
from __future__ import annotations

import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from oracle import operators as oops  # noqa: E402
from oracle import ref_shim  # noqa: E402
from oracle.make_golden import build_oracle_op, build_reference_op  # noqa: E402
from oracle.tiny_latent_net import TinyLatentCore  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

CASES = {
    "identity": dict(shape=(3, 16, 16), steps=6, batch=(), R=2, op=("identity",), sigma=0.05, omega=0.1,
                     gamma=1.0, eta=1.0),
    "blur9_batch": dict(shape=(3, 16, 16), steps=5, batch=(2,), R=1, op=("gblur", 9, 1.5), sigma=0.05,
                        omega=0.2, gamma=0.5, eta=0.5),
    "box2": dict(shape=(3, 16, 16), steps=5, batch=(), R=2, op=("box", 2), sigma=0.05, omega=0.1, gamma=1.0,
                 eta=0.0),
}


def make_reference_latent_net(ref, core, acp, steps):
    class Net(ref.networks.LatentEpsilonNetwork):
        def __init__(self):
            super().__init__(alphas_cumprod=acp)
            self.core = core
            self.calls = []

        def forward(self, x, t):
            e = self.core.eps(x, int(t))
            self.calls.append({"z_t": x.detach().clone(), "t": int(t), "eps": e.detach().clone()})
            return e

        @classmethod
        def from_pretrained(cls, *a, **k):
            raise NotImplementedError

        def set_sampling_parameters(self, num_sampling_steps, batch_size=1, num_reconstructions=1):
            self._batch_size, self._num_sampling_steps = batch_size, num_sampling_steps
            self._num_reconstructions = num_reconstructions
            ratio = 1000 // num_sampling_steps
            self.register_buffer("timesteps", torch.arange(0, num_sampling_steps, dtype=torch.int64) * ratio)

        @property
        def is_condition_initialized(self):
            return True

        def get_latent_shape(self, x_shape):
            return (4, x_shape[1] // 2, x_shape[2] // 2)

        def _decode(self, z, *, differentiable=False):
            return self.core.decode(z)

        def _encode(self, x, *, differentiable=False):
            return self.core.encode(x)

    return Net()


def run_case(name, cfg):
    ref = ref_shim.load_reference()
    from oracle.schedule import ddpm_linear_alphas_cumprod, padded_clipped_acp
    shape = tuple(cfg["shape"])
    core = TinyLatentCore(channels=shape[0])
    acp = torch.cat([torch.ones(1), ddpm_linear_alphas_cumprod()])
    net = make_reference_latent_net(ref, core, acp, cfg["steps"])
    oracle_op, extra = build_oracle_op(cfg["op"], shape)

    class _Wrapped(ref.operators.LinearOperator):
        def apply(self, x):
            lead = x.shape[: -len(shape)]
            out = oracle_op.apply(x.reshape(-1, *shape))
            return out.reshape(*lead, *out.shape[1:])

        def apply_transpose(self, yy):
            lead = yy.shape[: -len(oracle_op.y_shape)]
            out = oracle_op.adjoint(yy.reshape(-1, *oracle_op.y_shape))
            return out.reshape(*lead, *out.shape[1:])

    ref_op = ref.operators.IdentityOperator(x_shape=shape) if cfg["op"][0] == "identity" else _Wrapped(x_shape=shape)
    noise = ref.noise.GaussianNoise(sigma=cfg["sigma"])
    x_true = torch.rand((*cfg["batch"], *shape), generator=torch.Generator().manual_seed(0)) * 2 - 1
    problem = ref.inverse_problem.InverseProblem.from_clean_data(
        x_true, operator=ref_op, noise=noise, rng=torch.Generator().manual_seed(1))
    gz = torch.Generator().manual_seed(2)
    draws = []

    def draw(shape_):
        t = ref_shim.REAL_RANDN(shape_, generator=gz)
        draws.append(t)
        return t

    grads = []
    sampler = ref.samplers.PSLDSampler(net)
    with ref_shim.injected_noise(draw), ref_shim.recorded_autograd(grads):
        out = sampler(problem, num_sampling_steps=cfg["steps"], num_reconstructions=cfg["R"], gamma=cfg["gamma"],
                      omega=cfg["omega"], eta=cfg["eta"], decode_output=True)
    K = cfg["steps"] - 2
    calls = net.calls
    assert len(calls) == K + 1 and len(grads) == K and len(draws) == K + 1
    net.set_sampling_parameters(cfg["steps"])
    ts = net.timesteps.clone()
    meta = dict(name=name, shape=shape, latent_shape=[4, shape[1] // 2, shape[2] // 2], steps=cfg["steps"],
                batch=list(cfg["batch"]), R=cfg["R"], L=int(calls[0]["z_t"].shape[0]), op=list(cfg["op"]),
                sigma=cfg["sigma"], omega=cfg["omega"], gamma=cfg["gamma"], eta=cfg["eta"],
                t=[c["t"] for c in calls[:K]], t_prev=[int(ts[i - 1]) for i in range(len(ts) - 1, 1, -1)],
                s=int(ts[0]), torch=torch.__version__)
    arrays = dict(
        meta=np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8),
        acp=net.alphas_cumprod.numpy(), timesteps=ts.numpy(), y=problem.observation.numpy(), x_true=x_true.numpy(),
        z_init=draws[0].detach().numpy(), z_t=torch.stack([c["z_t"] for c in calls[:K]]).numpy(),
        eps=torch.stack([c["eps"] for c in calls[:K]]).numpy(), noise=torch.stack([d.detach() for d in draws[1:]]).numpy(),
        z_next=torch.stack([c["z_t"] for c in calls[1:]]).numpy(), grad=torch.stack(grads).numpy(),
        x_out=out.numpy(), **{f"net.{k}": v.numpy() for k, v in core.state_dict().items()}, **extra)
    np.savez_compressed(os.path.join(OUT, f"psld_{name}.npz"), **arrays)
    return meta


if __name__ == "__main__":
    torch.set_num_threads(1)
    for name, cfg in CASES.items():
        m = run_case(name, cfg)
        print("wrote psld", name, "L=", m["L"], "K=", len(m["t"]))
