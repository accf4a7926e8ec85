"""Record PGDM golden vectors from the UNMODIFIED reference PGDMSampler (build container only).
Usage: python -m oracle.make_golden_pgdm   ->  tests/golden/pgdm_<case>.npz"""
from __future__ import annotations

import importlib
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from oracle import ref_shim  # noqa: E402
from oracle.make_golden import build_oracle_op  # noqa: E402
from oracle.schedule import ddpm_linear_alphas_cumprod  # noqa: E402
from oracle.tiny_net import TinyEpsNet  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

CASES = {
    "identity": dict(shape=(3, 16, 16), steps=7, batch=(2,), R=2, op=("identity",), gw=0.3, eta=1.0),
    "inpaint": dict(shape=(3, 16, 16), steps=6, batch=(), R=2, op=("mask", 0.6, 9), gw=0.5, eta=0.5),
    "box4": dict(shape=(3, 16, 16), steps=6, batch=(), R=1, op=("box", 4), gw=0.2, eta=0.0),
}


def run_case(name, cfg):
    ref = ref_shim.load_reference()
    pgdm_mod = importlib.import_module("samplers.samplers.pgdm")
    shape = tuple(cfg["shape"])
    core = TinyEpsNet(channels=shape[0])
    rec = ref_shim.RecordingNet(core)
    pipe = ref_shim.FakeDDPMPipeline(rec, ddpm_linear_alphas_cumprod(), 1000)
    network = ref.networks.DDPMNetwork(pipe)
    oracle_op, extra = build_oracle_op(cfg["op"], shape)
    kind = cfg["op"][0]
    if kind == "identity":
        ref_op = ref.operators.IdentityOperator(x_shape=shape)
    elif kind == "mask":
        # the flat (m,) observation of the reference's InpaintingOperator breaks repeat_observation
        # (SURVEY App. B-7), so PGDM is recorded with the dense oracle operator under the reference ABC
        ref_op = None
    else:
        ref_op = None
    if ref_op is None:
        class _Wrapped(ref.operators.LinearOperator):
            def apply(self, x):
                lead = x.shape[: -len(shape)]
                out = dense.apply(x.reshape(-1, *shape))
                return out.reshape(*lead, *out.shape[1:])

            def apply_pseudo_inverse(self, yy):
                lead = yy.shape[: -len(dense.y_shape)]
                out = dense.pinv(yy.reshape(-1, *dense.y_shape))
                return out.reshape(*lead, *out.shape[1:])

        if kind == "mask":
            keep = (~torch.from_numpy(extra["mask"])).float()

            class _Dense:  # dense form of the mask: y has the shape of x, zeros at missing pixels
                x_shape = y_shape = shape

                @staticmethod
                def apply(x):
                    return x * keep

                pinv = apply
            dense = _Dense
        else:
            dense = oracle_op
        ref_op = _Wrapped(x_shape=shape)
    noise = ref.noise.GaussianNoise(sigma=0.05)
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
    sampler = pgdm_mod.PGDMSampler(network)
    with ref_shim.injected_noise(draw), ref_shim.recorded_autograd(grads):
        out = sampler(problem, num_sampling_steps=cfg["steps"], num_reconstructions=cfg["R"],
                      guidance_weight=cfg["gw"], eta=cfg["eta"], keep_reconstruction_dim=True)
    K = cfg["steps"] - 2
    calls = rec.calls
    assert len(calls) == K + 1 and len(grads) == K and len(draws) == K + 1
    network.set_sampling_parameters(cfg["steps"])
    ts = network.timesteps.clone()
    meta = dict(name=name, shape=shape, steps=cfg["steps"], batch=list(cfg["batch"]), R=cfg["R"],
                L=int(calls[0]["x_t"].shape[0]), op=list(cfg["op"]), gw=cfg["gw"], eta=cfg["eta"],
                t=[c["t"] for c in calls[:K]], t_prev=[int(ts[i - 1]) for i in range(len(ts) - 1, 1, -1)],
                s=int(ts[0]), torch=torch.__version__)
    arrays = dict(
        meta=np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8),
        acp=network.alphas_cumprod.numpy(), timesteps=ts.numpy(), y=problem.observation.numpy(),
        x_true=x_true.numpy(), x_init=draws[0].numpy(),
        x_t=torch.stack([c["x_t"] for c in calls[:K]]).numpy(), eps=torch.stack([c["eps"] for c in calls[:K]]).numpy(),
        z=torch.stack(draws[1:]).numpy(), x_next=torch.stack([c["x_t"] for c in calls[1:]]).numpy(),
        grad=torch.stack(grads).numpy(), x0_final=out.numpy(),
        **{f"net.{k}": v.numpy() for k, v in core.state_dict().items()}, **extra)
    np.savez_compressed(os.path.join(OUT, f"pgdm_{name}.npz"), **arrays)
    return meta


if __name__ == "__main__":
    torch.set_num_threads(1)
    for name, cfg in CASES.items():
        m = run_case(name, cfg)
        print("wrote pgdm", name, "L=", m["L"], "K=", len(m["t"]))
