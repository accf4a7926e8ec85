"""Record ReSample golden vectors from the UNMODIFIED reference ReSampleSampler (build container only).
Usage: python -m oracle.make_golden_resample   ->  tests/golden/resample_<case>.npz"""
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
from oracle.make_golden_psld import make_reference_latent_net  # noqa: E402
from oracle.tiny_latent_net import TinyLatentCore  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

CASES = {
    "identity": dict(shape=(3, 16, 16), steps=12, batch=(), R=2, op=("identity",), noise=("gaussian", 0.05),
                     kw=dict(sigma_scale=40.0, max_optimization_iters=6, eta=1.0, inter_timesteps=2,
                             time_travel_interval=2, stage_splits=3)),
    "blur9_poisson": dict(shape=(3, 16, 16), steps=10, batch=(2,), R=1, op=("gblur", 9, 1.5), noise=("poisson", 4.0),
                          kw=dict(sigma_scale=10.0, max_optimization_iters=4, eta=0.5, inter_timesteps=3,
                                  time_travel_interval=3, stage_splits=2)),
    # the reference's own InpaintingOperator (flatten=True): nn.MSELoss averages over the m KEPT pixels, which scales
    # the optimisers' gradients and the `< eps^2` stop rule (resample_kernels.py:43-51,71-91)
    "inpaint": dict(shape=(3, 16, 16), steps=10, batch=(), R=1, op=("mask", 0.7, 7), noise=("gaussian", 0.05),
                    kw=dict(sigma_scale=40.0, max_optimization_iters=8, eta=1.0, inter_timesteps=2,
                            time_travel_interval=2, stage_splits=3)),
}


def run_case(name, cfg):
    ref = ref_shim.load_reference()
    resample_mod = importlib.import_module("samplers.samplers.resample")
    from oracle.schedule import ddpm_linear_alphas_cumprod
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

    if cfg["op"][0] == "identity":
        ref_op = ref.operators.IdentityOperator(x_shape=shape)
    elif cfg["op"][0] == "mask":
        ref_op = ref.operators.InpaintingOperator(x_shape=shape, mask=torch.from_numpy(extra["mask"]))
    else:
        ref_op = _Wrapped(x_shape=shape)
    nk, nparam = cfg["noise"]
    noise = ref.noise.GaussianNoise(sigma=nparam) if nk == "gaussian" else ref.noise.PoissonNoise(rate=nparam)
    x_true = torch.rand((*cfg["batch"], *shape), generator=torch.Generator().manual_seed(0)) * 2 - 1
    problem = ref.inverse_problem.InverseProblem.from_clean_data(
        x_true, operator=ref_op, noise=noise, rng=torch.Generator().manual_seed(1))
    gz = torch.Generator().manual_seed(2)
    draws = []

    def draw(shape_):
        t = ref_shim.REAL_RANDN(shape_, generator=gz)
        draws.append(t)
        return t

    sampler = resample_mod.ReSampleSampler(net)
    with ref_shim.injected_noise(draw):
        out = sampler(problem, num_sampling_steps=cfg["steps"], num_reconstructions=cfg["R"], decode_output=True,
                      **cfg["kw"])
    net.set_sampling_parameters(cfg["steps"])
    meta = dict(name=name, shape=shape, latent_shape=[4, shape[1] // 2, shape[2] // 2], steps=cfg["steps"],
                batch=list(cfg["batch"]), R=cfg["R"], L=int(draws[0].shape[0]), op=list(cfg["op"]),
                noise=[nk, nparam], kw=cfg["kw"], n_draws=len(draws), n_net_calls=len(net.calls),
                torch=torch.__version__)
    arrays = dict(
        meta=np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8),
        acp=net.alphas_cumprod.numpy(), timesteps=net.timesteps.numpy(), y=problem.observation.numpy(),
        x_true=x_true.numpy(), x_out=out.detach().numpy(),
        draws=torch.stack([d.detach() for d in draws]).numpy(),
        call_z=torch.stack([c["z_t"] for c in net.calls]).numpy(), call_t=np.array([c["t"] for c in net.calls]),
        **{f"net.{k}": v.numpy() for k, v in core.state_dict().items()}, **extra)
    np.savez_compressed(os.path.join(OUT, f"resample_{name}.npz"), **arrays)
    return meta


if __name__ == "__main__":
    torch.set_num_threads(1)
    for name, cfg in CASES.items():
        m = run_case(name, cfg)
        print("wrote resample", name, "L=", m["L"], "draws=", m["n_draws"], "net calls=", m["n_net_calls"])
