#!/usr/bin/env python
"""Development stress of blur_k1_tc: 300 repeats per batch size (one-plane and persistent instantiations, 75-300 planes)
with a varying cache state in between; every repeat must be bit-identical to the first (a phase / parity slip of the
persistent mbarrier protocol would show as a mismatch or a trap)."""
import sys, torch
sys.path.insert(0, '/root/repo')
from samplers_b200 import _native, operators as pops
dev = torch.device("cuda:0")
op = pops.GaussianBlurOperator((3, 256, 256), 61, 3.0).to(dev)
nat = op._native_cached(dev)
n = nat.n
g = torch.Generator(device=dev).manual_seed(1)
for L, rep in ((64, 64), (30, 1), (50, 5), (25, 25), (100, 4)):
    x = torch.randn(L, n, device=dev, generator=g); e = torch.randn(L, n, device=dev, generator=g)
    y = torch.randn(L // rep, n, device=dev, generator=g)
    ws = torch.empty(nat.workspace_bytes(L) // 4, device=dev)
    ref = None; bad = 0
    big = torch.empty(64 << 20, device=dev)
    for it in range(300):
        cot = torch.full_like(x, float("nan")); part = torch.full((L, nat.err_parts), float("nan"), device=dev)
        if it % 3 == 0: big.fill_(it)          # vary the cache / timing state
        _native.dps_pre(nat, x, e, y, rep, 0.9, 0.43, 25.0, cot, part, ws)
        if ref is None: ref = (cot.clone(), part.clone())
        elif not (torch.equal(cot, ref[0]) and torch.equal(part, ref[1])): bad += 1
    torch.cuda.synchronize()
    print(L, rep, "mismatching repeats:", bad, "finite:", bool(torch.isfinite(ref[0]).all()), flush=True)
