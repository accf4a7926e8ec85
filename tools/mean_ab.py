#!/usr/bin/env python
"""A/B of the bridge-mean role of blur_k1_tc at config 2 (L = 16): K1 / K2 over a ring of cold buffer sets (the ring
pass replayed as one CUDA graph, as bench.py's by_config does) for the classic pair and for the pair with the mean, the
latter at several start lags of the mean CTAs (PSX_MEAN_LAG_NS)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from samplers_b200 import _native, operators as pops  # noqa: E402

dev = torch.device("cuda:0")
L = int(sys.argv[1]) if len(sys.argv) > 1 else 16
op = pops.GaussianBlurOperator((3, 256, 256), 61, 3.0).to(dev)
nat = op._native_cached(dev)
n = nat.n
gen = torch.Generator(device=dev).manual_seed(0)
nsets = max(2, -(-4 * 126 * 2 ** 20 // (8 * L * n * 4)))
S = []
for _ in range(nsets):
    d = {k: torch.randn(L, n, device=dev, generator=gen) for k in ("x", "eps", "v", "z")}
    d.update(cot=torch.empty(L, n, device=dev), out=torch.empty(L, n, device=dev), mean=torch.empty(L, n, device=dev),
             part=torch.empty(L, nat.err_parts, device=dev), ws=torch.empty(max(nat.workspace_bytes(L) // 4, 1), device=dev))
    S.append(d)
y = torch.randn(1, nat.n_y, device=dev, generator=gen)


def k1(i):
    d = S[i]
    _native.dps_pre(nat, d["x"], d["eps"], y, L, 0.8, 0.6, 400.0, d["cot"], d["part"], d["ws"])


def k2(i):
    d = S[i]
    _native.dps_post(d["x"], d["eps"], d["cot"], d["v"], d["z"], d["part"], nat.err_parts, n, 0.8, 0.6, 0.99, 0.01,
                     0.05, 1.0, d["out"], None)


def k1m(i):
    d = S[i]
    _native.dps_pre_mean(nat, d["x"], d["eps"], y, L, 0.8, 0.6, 400.0, 0.99, 0.01, d["cot"], d["part"], d["mean"], d["ws"],
                         **({} if os.environ.get("NO_Z") else dict(z=d["z"], std=0.05)))


def k2m(i):
    d = S[i]
    _native.dps_post_mean(d["mean"], d["cot"], d["v"], d["z"] if os.environ.get("NO_Z") else None, d["part"], nat.err_parts, n,
                          0.6, 0.05 if os.environ.get("NO_Z") else 0.0, 1.0, d["out"], None)


if os.environ.get("ONCE"):  # one launch of each kernel of the pair (for ncu)
    for i in range(nsets):
        k1m(i); k2m(i)
    torch.cuda.synchronize()
    sys.exit(0)

for rep in range(2):
    print(f"classic: K1 {bench._rotate_time(k1, nsets, 10) * 1e3:.2f} us  K2 {bench._rotate_time(k2, nsets, 10) * 1e3:.2f} us")
    for lag in [int(v) for v in os.environ.get("LAGS", "0,2000,6000,8000,9000,10000,11000,12000,14000").split(",")]:
        os.environ["PSX_MEAN_LAG_NS"] = str(lag)
        _native.load().psx_reload_env()
        print(f"mean lag {lag:5d} ns: K1 {bench._rotate_time(k1m, nsets, 10) * 1e3:.2f} us  "
              f"K2 {bench._rotate_time(k2m, nsets, 10) * 1e3:.2f} us")
