"""Record golden vectors from the UNMODIFIED reference (run in the build
container only: needs /root/reference).  Usage:  python -m oracle.make_golden

For every case the reference's own ``DPSSampler`` (samplers/samplers/dps.py)
is run on CPU fp32 with injected noise tensors; each network call and each
``torch.autograd.grad`` result is recorded.  Operators that the reference lacks
(blur / motion / box) are supplied as subclasses of the reference's
``LinearOperator`` whose ``apply`` is the oracle operator, so the gradient is
produced by the reference's autograd path, not by our adjoint.

Output: tests/golden/<case>.npz  (small; committed).
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from oracle import dps as odps  # noqa: E402
from oracle import operators as oops  # noqa: E402
from oracle import ref_shim  # noqa: E402
from oracle.schedule import ddpm_linear_alphas_cumprod  # noqa: E402
from oracle.tiny_net import TinyEpsNet  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

CASES = {
    # name: dict(shape, steps, batch_shape, R, op, noise, gamma, eta, schedule)
    "identity_gauss": dict(shape=(3, 16, 16), steps=8, batch=(), R=2, op=("identity",),
                           noise=("gaussian", 0.05), gamma=1.0, eta=1.0, schedule="ddpm"),
    "identity_poisson": dict(shape=(3, 16, 16), steps=6, batch=(), R=1, op=("identity",),
                             noise=("poisson", 4.0), gamma=0.7, eta=1.0, schedule="ddpm"),
    "identity_batch": dict(shape=(3, 16, 16), steps=6, batch=(3,), R=1, op=("identity",),
                           noise=("gaussian", 0.1), gamma=1.0, eta=1.0, schedule="ddpm"),
    "identity_mock_eta05": dict(shape=(3, 8, 8), steps=5, batch=(), R=2, op=("identity",),
                                noise=("gaussian", 0.2), gamma=0.5, eta=0.5, schedule="mock"),
    "identity_ddim": dict(shape=(3, 16, 16), steps=6, batch=(), R=1, op=("identity",),
                          noise=("gaussian", 0.05), gamma=1.0, eta=0.0, schedule="ddpm"),
    "inpaint_gauss": dict(shape=(3, 16, 16), steps=8, batch=(), R=2, op=("mask", 0.7, 7),
                          noise=("gaussian", 0.05), gamma=1.0, eta=1.0, schedule="ddpm"),
    "blur9_gauss": dict(shape=(3, 16, 24), steps=6, batch=(), R=2, op=("gblur", 9, 1.5),
                        noise=("gaussian", 0.05), gamma=1.0, eta=1.0, schedule="ddpm"),
    "blur61_gauss": dict(shape=(3, 64, 64), steps=4, batch=(), R=1, op=("gblur", 61, 3.0),
                         noise=("gaussian", 0.05), gamma=1.0, eta=1.0, schedule="ddpm"),
    "motion9_gauss": dict(shape=(3, 16, 16), steps=6, batch=(), R=1, op=("motion", 9, 30.0),
                          noise=("gaussian", 0.05), gamma=1.0, eta=1.0, schedule="ddpm"),
    "box4_gauss": dict(shape=(3, 16, 16), steps=6, batch=(), R=2, op=("box", 4),
                       noise=("gaussian", 0.05), gamma=1.0, eta=1.0, schedule="ddpm"),
}


def build_oracle_op(spec, shape):
    kind = spec[0]
    if kind == "identity":
        return oops.OracleIdentity(shape), {}
    if kind == "mask":
        g = torch.Generator().manual_seed(spec[2])
        mask = torch.rand(shape, generator=g) < spec[1]  # True = missing
        return oops.OracleMaskGather(shape, mask), {"mask": mask.numpy()}
    if kind == "gblur":
        op = oops.OracleGaussianBlur(shape, spec[1], spec[2])
        return op, {"taps": op.taps_h.numpy()}
    if kind == "motion":
        k2d = oops.motion_line_kernel(spec[1], spec[2])
        return oops.OracleConv2dBlur(shape, k2d), {"kernel2d": k2d.numpy()}
    if kind == "box":
        return oops.OracleBoxDownsample(shape, spec[1]), {}
    raise ValueError(kind)


def build_reference_op(ref, spec, shape, oracle_op, extra):
    rops = ref.operators
    kind = spec[0]
    if kind == "identity":
        return rops.IdentityOperator(x_shape=shape)
    if kind == "mask":
        return rops.InpaintingOperator(x_shape=shape, mask=torch.from_numpy(extra["mask"]))

    class _Wrapped(rops.LinearOperator):  # reference ABC, oracle arithmetic
        def apply(self, x):
            lead = x.shape[: -len(shape)]
            out = oracle_op.apply(x.reshape(-1, *shape))
            return out.reshape(*lead, *out.shape[1:])

    return _Wrapped(x_shape=shape)


def run_case(name, cfg):
    ref = ref_shim.load_reference()
    shape = tuple(cfg["shape"])
    torch.manual_seed(0)
    net_core = TinyEpsNet(channels=shape[0])
    rec = ref_shim.RecordingNet(net_core)

    if cfg["schedule"] == "ddpm":
        pipe = ref_shim.FakeDDPMPipeline(rec, ddpm_linear_alphas_cumprod(), 1000)
    else:  # tests/samplers/test_pgdm.py:62-66,86 -- s = timesteps[0] = 1 != 0
        pipe = ref_shim.FakeDDPMPipeline(
            rec, torch.linspace(0.99, 0.5, cfg["steps"], dtype=torch.float32), cfg["steps"],
            custom_ascending=lambda n: list(range(1, n + 1)))
    network = ref.networks.DDPMNetwork(pipe)

    oracle_op, extra = build_oracle_op(cfg["op"], shape)
    ref_op = build_reference_op(ref, cfg["op"], shape, oracle_op, extra)
    nk, nparam = cfg["noise"]
    noise = ref.noise.GaussianNoise(sigma=nparam) if nk == "gaussian" else ref.noise.PoissonNoise(rate=nparam)

    gx = torch.Generator().manual_seed(0)
    x_true = torch.rand((*cfg["batch"], *shape), generator=gx) * 2 - 1
    problem = ref.inverse_problem.InverseProblem.from_clean_data(
        x_true, operator=ref_op, noise=noise, rng=torch.Generator().manual_seed(1))
    y = problem.observation

    gz = torch.Generator().manual_seed(2)
    draws = []

    def draw(shape_):
        t = ref_shim.REAL_RANDN(shape_, generator=gz)
        draws.append(t)
        return t

    grads = []
    sampler = ref.samplers.DPSSampler(network)
    with ref_shim.injected_noise(draw), ref_shim.recorded_autograd(grads):
        x0_final = sampler(problem, num_sampling_steps=cfg["steps"], num_reconstructions=cfg["R"],
                           gamma=cfg["gamma"], eta=cfg["eta"], keep_reconstruction_dim=True)

    K = cfg["steps"] - 2
    calls = rec.calls
    assert len(calls) == K + 1 and len(grads) == K and len(draws) == K + 1, (len(calls), len(grads), len(draws))
    network.set_sampling_parameters(cfg["steps"])
    acp = network.alphas_cumprod.clone()
    ts = network.timesteps.clone()

    x_t = torch.stack([c["x_t"] for c in calls[:K]])
    eps = torch.stack([c["eps"] for c in calls[:K]])
    x_next = torch.stack([c["x_t"] for c in calls[1:]])
    z = torch.stack(draws[1:])
    grad = torch.stack(grads)
    t_list = [c["t"] for c in calls[:K]]
    tp_list = [int(ts[i - 1]) for i in range(len(ts) - 1, 1, -1)]
    assert t_list == [int(ts[i]) for i in range(len(ts) - 1, 1, -1)]

    # oracle closed form, teacher-forced on the reference's states: d, v, e2 per step
    L = x_t.shape[1]
    y_flat = y if len(cfg["batch"]) else y.unsqueeze(0)
    w = odps.likelihood_weight(nk, torch.tensor(nparam))
    d_all, v_all, e2_all = [], [], []
    for k in range(K):
        xt = x_t[k].clone().requires_grad_()
        e = net_core(xt, t_list[k])
        d, e2, _ = odps.k1_reference(xt.detach(), e.detach(), acp_t=acp[t_list[k]], op=oracle_op,
                                     y=y_flat, weight=w)
        (v,) = torch.autograd.grad(e, xt, grad_outputs=d)
        d_all.append(d); v_all.append(v); e2_all.append(e2)

    meta = dict(name=name, shape=shape, steps=cfg["steps"], batch=list(cfg["batch"]), R=cfg["R"], L=L,
                op=list(cfg["op"]), noise=[nk, nparam], gamma=cfg["gamma"], eta=cfg["eta"],
                schedule=cfg["schedule"], t=t_list, t_prev=tp_list, s=int(ts[0]),
                torch=torch.__version__)
    arrays = dict(
        meta=np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8),
        acp=acp.numpy(), timesteps=ts.numpy(), y=y.numpy(), x_true=x_true.numpy(),
        x_init=draws[0].numpy(), x_t=x_t.numpy(), eps=eps.numpy(), z=z.numpy(),
        x_next=x_next.numpy(), grad=grad.numpy(),
        final_eps=calls[K]["eps"].numpy(), x0_final=x0_final.numpy(),
        d=torch.stack(d_all).numpy(), v=torch.stack(v_all).numpy(), e2=torch.stack(e2_all).numpy(),
        **{f"net.{k}": v.numpy() for k, v in net_core.state_dict().items()},
        **extra,
    )
    os.makedirs(OUT, exist_ok=True)
    np.savez_compressed(os.path.join(OUT, f"dps_{name}.npz"), **arrays)
    return meta


if __name__ == "__main__":
    torch.set_num_threads(1)  # deterministic reduction order for the recording
    for name, cfg in CASES.items():
        m = run_case(name, cfg)
        print("wrote", name, "L=", m["L"], "K=", len(m["t"]))
