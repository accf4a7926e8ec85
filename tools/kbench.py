#!/usr/bin/env python
"""Kernel-only micro-benchmark of the fused DPS step (K1 + K2) through the C ABI.

Two timing modes, both with CUDA events on the launching stream:
  rotate (default): launches run back to back over a ring of independent buffer sets whose total
      footprint is >= 4x the 126 MB L2, one event pair around the whole ring pass -- every launch sees
      cold inputs, and the ~4 us event/launch latency of a single tiny kernel is amortised
      (this is how the kernels run inside the sampler: queued behind other work);
  flush: one event pair per launch, a 512 MB write between launches (upper bound: includes the
      launch latency of an idle stream).
Prints one JSON line per case: algorithmic GB/s = (16 [K1] / 24 [K2] B per element) / time and the
fraction of the measured HBM peak.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from samplers_b200 import _native, operators as pops  # noqa: E402

SHAPE = (3, 256, 256)  # --size changes H = W
BF16_OPS = ("identity", "mask", "box4", "gblur61", "gblur13")  # operators with a bf16-state K1


def peak():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        return 6650.0


def make_op(kind):
    if kind == "identity":
        return pops.IdentityOperator(SHAPE)
    if kind == "mask":
        return pops.RandomInpaintingOperator(SHAPE, 0.7, flatten=False)
    if kind == "box4":
        return pops.BoxDownsampleOperator(SHAPE, 4)
    if kind == "gblur61":
        return pops.GaussianBlurOperator(SHAPE, 61, 3.0)
    if kind == "gblur13":  # 13 taps -> the K = 16 instantiation: same structure, 0.4x the FMA work
        return pops.GaussianBlurOperator(SHAPE, 13, 1.5)
    if kind == "motion61":
        return pops.MotionBlurOperator(SHAPE, kernel_size=61, angle_deg=30.0)
    if kind == "motion61s":  # steep line: 61 segments of 1-2 chunks
        return pops.MotionBlurOperator(SHAPE, kernel_size=61, angle_deg=80.0)
    if kind == "walk61":     # random camera-shake trajectory
        return pops.MotionBlurOperator(SHAPE, kernel_size=61, intensity=0.5, seed=0)
    raise ValueError(kind)


def time_flush(fn, iters, flush):
    ts = []
    for _ in range(iters):
        flush.fill_(1.0)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(0); e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    return statistics.median(ts)


def time_rotate(fn, nsets, passes):
    ts = []
    for _ in range(passes):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        s.record()
        for i in range(nsets):
            fn(i)
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e) / nsets)
    return statistics.median(ts)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ops", default="identity,mask,box4,gblur61")
    ap.add_argument("--batches", default="16,64")
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--mode", default="rotate", choices=["rotate", "flush"])
    ap.add_argument("--size", type=int, default=256)
    args = ap.parse_args()
    global SHAPE
    SHAPE = (3, args.size, args.size)
    dev = torch.device("cuda:0")
    pk = peak()
    flush = torch.empty(128 * 1024 * 1024 if args.mode == "flush" else 1, device=dev)
    gen = torch.Generator(device=dev).manual_seed(0)
    for kind in args.ops.split(","):
        op = make_op(kind).to(dev)
        nat = op._native_cached(dev)
        for L in [int(b) for b in args.batches.split(",")]:
            n = nat.n
            per_set = 7 * L * n * 4  # x, eps, v, z, cot, out, ws
            nsets = max(2, -(-4 * 126 * 2**20 // per_set)) if args.mode == "rotate" else 1
            S = []
            for _ in range(nsets):
                d = dict(x=torch.randn(L, n, device=dev, generator=gen), eps=torch.randn(L, n, device=dev, generator=gen),
                         v=torch.randn(L, n, device=dev, generator=gen), z=torch.randn(L, n, device=dev, generator=gen),
                         cot=torch.empty(L, n, device=dev), out=torch.empty(L, n, device=dev),
                         part=torch.empty(L, nat.err_parts, device=dev))
                wsb = nat.workspace_bytes(L)
                d["ws"] = torch.empty(wsb // 4, device=dev) if wsb else None
                if kind in BF16_OPS:
                    for a, b in (("xh", "x"), ("eh", "eps"), ("vh", "v"), ("ch", "cot"), ("oh", "out")):
                        d[a] = d[b].to(torch.bfloat16)
                S.append(d)
            y = torch.randn(1, nat.n_y, device=dev, generator=gen)

            def k1(i):
                d = S[i]
                _native.dps_pre(nat, d["x"], d["eps"], y, L, 0.8, 0.6, 400.0, d["cot"], d["part"], d["ws"])

            def k2(i):
                d = S[i]
                _native.dps_post(d["x"], d["eps"], d["cot"], d["v"], d["z"], d["part"], nat.err_parts, n, 0.8, 0.6,
                                 0.99, 0.01, 0.05, 1.0, d["out"], None)

            def k2p(i):  # noise drawn inside the kernel (20 B/elem of traffic)
                d = S[i]
                _native.dps_post_philox(d["x"], d["eps"], d["cot"], d["v"], d["part"], nat.err_parts, n, 0.8, 0.6,
                                        0.99, 0.01, 0.05, 1.0, 1234, i, d["out"], None)

            def k12h(i):  # bf16 state: K1 + K2 with in-kernel noise (18 B/elem); identity / mask only
                d = S[i]
                _native.dps_pre_bf16(nat, d["xh"], d["eh"], y, L, 0.8, 0.6, 400.0, d["ch"], d["part"], ws=d["ws"])
                _native.dps_post_bf16(d["xh"], d["eh"], d["ch"], d["vh"], None, d["part"], nat.err_parts, n, 0.8, 0.6,
                                      0.99, 0.01, 0.05, 1.0, d["oh"], None, philox=(1234, i))

            def zgen(i):  # what the Philox K2 replaces: the generator kernel that writes z
                S[i]["z"].normal_()

            for i in range(nsets):
                k1(i); k2(i); k2p(i)
            if args.mode == "rotate":
                m1, m2 = time_rotate(k1, nsets, args.iters), time_rotate(k2, nsets, args.iters)
                m2p, mz = time_rotate(k2p, nsets, args.iters), time_rotate(zgen, nsets, args.iters)
            else:
                m1, m2 = time_flush(k1, args.iters, flush), time_flush(k2, args.iters, flush)
                m2p, mz = time_flush(k2p, args.iters, flush), time_flush(zgen, args.iters, flush)
            mh = None
            if kind in BF16_OPS:
                for i in range(nsets):
                    k12h(i)
                mh = time_rotate(k12h, nsets, args.iters) if args.mode == "rotate" else time_flush(k12h, args.iters, flush)
            b1, b2 = 16 * L * n, 24 * L * n
            print(json.dumps({
                "op": kind, "size": args.size, "L": L, "mode": args.mode, "nsets": nsets, "k1_us": m1 * 1e3, "k2_us": m2 * 1e3,
                "k1_gbs": b1 / m1 / 1e6, "k2_gbs": b2 / m2 / 1e6,
                "fused_gbs": (b1 + b2) / (m1 + m2) / 1e6, "fused_frac": (b1 + b2) / (m1 + m2) / 1e6 / pk,
                "k1_frac": b1 / m1 / 1e6 / pk, "k2_frac": b2 / m2 / 1e6 / pk, "peak": pk,
                "k2_philox_us": m2p * 1e3, "torch_normal_us": mz * 1e3,
                **({"bf16_step_us": mh * 1e3, "bf16_step_gbs": 18 * L * n / mh / 1e6} if mh else {})}), flush=True)
            del S


if __name__ == "__main__":
    main()
