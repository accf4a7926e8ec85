#!/usr/bin/env python
"""Development check of the tensor-core blur K1 (psx_tcblur.cu) through the C ABI: the same call with PSX_NO_TC=1
(CUDA-core strip kernels) and without, both against an fp64 torch evaluation of the same operator.  Prints one JSON
line per case; exits non-zero when an error exceeds 1e-5."""
from __future__ import annotations

import json
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from samplers_b200 import _native, operators as pops  # noqa: E402


def ref64(op, x, eps, y, sa, s1, w, obs_repeat):
    L = x.shape[0]
    C, H, W = op.x_shape
    k = op.taps_h.double().to(x.device)
    x0 = ((x.double() - s1 * eps.double()) / sa).view(L * C, 1, H, W)
    R = k.numel() // 2

    def rows(t, kk):
        return F.conv2d(t, kk.view(1, 1, 1, -1), padding=(0, R))

    def cols(t, kk):
        return F.conv2d(t, kk.view(1, 1, -1, 1), padding=(R, 0))

    ax = cols(rows(x0, k), k)
    yy = y.double().view(-1, C, H, W).repeat_interleave(obs_repeat, 0).view(L * C, 1, H, W)
    r = yy - ax
    err = (r * r).view(L, -1).sum(1)
    kf = k.flip(0)
    cot = rows(cols(r, kf), kf) * (w / sa)
    return cot.view(L, -1), err


def run(nat, x, eps, y, L, sa, s1, w, obs_repeat):
    cot = torch.full_like(x, float("nan"))
    part = torch.full((L, nat.err_parts), float("nan"), device=x.device)
    wsb = nat.workspace_bytes(L)
    ws = torch.empty(wsb // 4, device=x.device) if wsb else None
    _native.dps_pre(nat, x, eps, y, obs_repeat, sa, s1, w, cot, part, ws)
    torch.cuda.synchronize()
    return cot, part.double().sum(1)


def main():
    dev = torch.device("cuda:0")
    gen = torch.Generator(device=dev).manual_seed(0)
    op = pops.GaussianBlurOperator((3, 256, 256), 61, 3.0).to(dev)
    nat = op._native_cached(dev)
    worst = 0.0
    # yscale None: the observation of a late timestep, y = A x0 + N(0, 0.05^2) (residual at the noise level);
    # sa = 0.00633: the first timestep of the 1000-step schedule (x0 and the residual are O(1 / sa))
    for (L, obs_repeat, sa, s1, w, yscale) in [(1, 1, 0.8, 0.6, 400.0, 1.0), (2, 2, 0.05, 0.998, 400.0, 1.0),
                                              (16, 16, 0.9, 0.43, 25.0, 1.0), (5, 1, 0.999, 0.03, 400.0, 1.0),
                                              (4, 4, 0.00633, 0.99998, 400.0, 1.0), (3, 1, 0.9995, 0.0316, 400.0, None),
                                              (2, 1, 1.0, 0.0, 1.0, None),
                                              # more planes than resident cluster pairs (74): the persistent loop --
                                              # 90 planes (some pairs run two planes, some one), 150 (two or three), 192
                                              (30, 1, 0.7, 0.71, 100.0, 1.0), (50, 5, 0.3, 0.95, 400.0, 1.0),
                                              (64, 64, 0.9, 0.43, 25.0, 1.0), (25, 1, 0.9995, 0.0316, 400.0, None)]:
        n = nat.n
        x = torch.randn(L, n, device=dev, generator=gen)
        eps = torch.randn(L, n, device=dev, generator=gen)
        if yscale is None:
            C, H, W = op.x_shape
            k = op.taps_h.double().to(dev)
            R = k.numel() // 2
            x0 = ((x.double() - s1 * eps.double()) / sa).view(L * C, 1, H, W)
            ax = F.conv2d(F.conv2d(x0, k.view(1, 1, 1, -1), padding=(0, R)), k.view(1, 1, -1, 1), padding=(R, 0))
            y = (ax.view(L, n) + 0.05 * torch.randn(L, n, device=dev, generator=gen, dtype=torch.float64)).float()
        else:
            y = torch.randn(L // obs_repeat, nat.n_y, device=dev, generator=gen) * yscale
        rc, re = ref64(op, x, eps, y, sa, s1, w, obs_repeat)
        out = {"L": L, "obs_repeat": obs_repeat, "sa": sa}
        for name, env in (("cuda_core", "1"), ("tc", None)):
            if env:
                os.environ["PSX_NO_TC"] = env
            else:
                os.environ.pop("PSX_NO_TC", None)
            _native.reload_env()
            cot, err = run(nat, x, eps, y, L, sa, s1, w, obs_repeat)
            e_cot = ((cot.double() - rc).norm() / rc.norm()).item()
            e_max = ((cot.double() - rc).abs().max() / rc.abs().max()).item()
            e_err = ((err - re).abs() / re).max().item()
            out[name] = {"cot_rel_fro": e_cot, "cot_rel_max": e_max, "err_rel": e_err}
            if name == "tc":
                worst = max(worst, e_cot, e_max, e_err)
        print(json.dumps(out), flush=True)
    # timing, rotate mode over cold buffers
    for L in (16, 32, 64):
        n = nat.n
        nsets = max(2, -(-4 * 126 * 2**20 // (4 * L * n * 4)))
        S = [dict(x=torch.randn(L, n, device=dev), eps=torch.randn(L, n, device=dev), cot=torch.empty(L, n, device=dev),
                  part=torch.empty(L, nat.err_parts, device=dev), ws=torch.empty(nat.workspace_bytes(L) // 4, device=dev))
             for _ in range(nsets)]
        y = torch.randn(1, nat.n_y, device=dev)
        res = {"L": L, "nsets": nsets}
        for name, env in (("cuda_core", "1"), ("tc", None), ("tc_persist", "one"), ("tc", None)):
            os.environ.pop("PSX_NO_TC", None)
            os.environ.pop("PSX_TC_PERSIST", None)
            if env == "1":
                os.environ["PSX_NO_TC"] = env
            elif env == "one":
                os.environ["PSX_TC_PERSIST"] = "1"
            _native.reload_env()

            def k1(i):
                d = S[i]
                _native.dps_pre(nat, d["x"], d["eps"], y, L, 0.8, 0.6, 400.0, d["cot"], d["part"], d["ws"])

            for i in range(nsets):
                k1(i)
            ts = []
            for _ in range(10):
                s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                torch.cuda.synchronize()
                s.record()
                for i in range(nsets):
                    k1(i)
                e.record()
                torch.cuda.synchronize()
                ts.append(s.elapsed_time(e) / nsets * 1e3)
            ts.sort()
            res.setdefault(name + "_k1_us", []).append(round(ts[len(ts) // 2], 2))
        print(json.dumps(res), flush=True)
    print("worst tc error", worst)
    return 0 if worst < 1e-5 else 1


if __name__ == "__main__":
    sys.exit(main())
