"""Eager vs CUDA-graph replay of the DPS timestep: ms per guided timestep, CUDA events, same run state.

    python tools/graph_bench.py [--config tiny|ddpm-celebahq-256] [--size 64] [--batch 1] [--op identity|blur] [--steps 30]
"""
import argparse
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="ddpm-celebahq-256")
    ap.add_argument("--size", type=int, default=64)
    ap.add_argument("--batch", type=int, default=1)
    ap.add_argument("--op", default="identity")
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--net-dtype", default="fp32", choices=["fp32", "bf16"])
    ap.add_argument("--state-dtype", default="fp32", choices=["fp32", "bf16"])
    ap.add_argument("--philox", type=int, default=0)
    a = ap.parse_args()
    from samplers_b200 import operators as P
    from samplers_b200.inverse_problem import InverseProblem
    from samplers_b200.networks import DDPMNetwork
    from samplers_b200.noise import GaussianNoise
    from samplers_b200.samplers import DPSSampler
    dev = "cuda:0"
    shape = (3, a.size, a.size)
    net = DDPMNetwork.from_config(a.config, device=dev, torch_dtype=torch.bfloat16 if a.net_dtype == "bf16" else None)
    op = (P.IdentityOperator(shape) if a.op == "identity" else P.GaussianBlurOperator(shape)).to(dev)
    gen = torch.Generator(device=dev).manual_seed(0)
    x = torch.rand(shape, device=dev, generator=gen) * 2 - 1
    y = op.apply(x[None])[0] + 0.05 * torch.randn(op.y_shape, device=dev, generator=gen)
    prob = InverseProblem(operator=op, observation=y, noise=GaussianNoise(sigma=0.05))
    out = {"config": a.config, "shape": list(shape), "batch": a.batch, "operator": a.op, "steps": a.steps,
           "net_dtype": a.net_dtype, "state_dtype": a.state_dtype, "philox": bool(a.philox)}
    for mode in ("eager", "graph"):
        s = DPSSampler(net, cuda_graph=mode == "graph", philox_seed=0 if a.philox else None,
                       state_dtype=torch.bfloat16 if a.state_dtype == "bf16" else None)
        run = s.prepare(prob, num_sampling_steps=1000, num_reconstructions=a.batch)
        try:
            if mode == "graph":
                t0 = time.perf_counter()
                run.capture()
                torch.cuda.synchronize()
                out["capture_s"] = round(time.perf_counter() - t0, 3)
            for k in range(a.warmup):
                run.step(k)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            w0 = time.perf_counter()
            e0.record()
            for k in range(a.warmup, a.warmup + a.steps):
                run.step(k)
            e1.record()
            host_issue = time.perf_counter() - w0
            torch.cuda.synchronize()
            out[mode] = {"ms_per_step": round(e0.elapsed_time(e1) / a.steps, 4),
                         "host_issue_ms_per_step": round(1e3 * host_issue / a.steps, 4),
                         "finite": bool(torch.isfinite(run.x).all())}
        finally:
            s.release()
    out["speedup"] = round(out["eager"]["ms_per_step"] / out["graph"]["ms_per_step"], 3)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
