#!/usr/bin/env python
"""Times the torch eps-network (ddpm-celebahq-256 UNet, random init) forward + input-VJP at batch 16, 256^2
under a few torch settings -- to pick the fastest *fp32* eager configuration for bench.py."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from samplers_b200.networks.unet2d import CELEBAHQ_256, UNet2DModel  # noqa: E402


def run(name, channels_last, matmul_tf32, autocast=None, batch=16, iters=5, param_dtype=None):
    torch.backends.cuda.matmul.allow_tf32 = matmul_tf32
    torch.backends.cudnn.benchmark = True
    torch.manual_seed(0)
    net = UNet2DModel(**CELEBAHQ_256).cuda().eval().requires_grad_(False)
    if param_dtype is not None:
        net = net.to(param_dtype)
    if channels_last:
        net = net.to(memory_format=torch.channels_last)
    x = torch.randn(batch, 3, 256, 256, device="cuda", dtype=param_dtype or torch.float32)
    g = torch.randn(batch, 3, 256, 256, device="cuda", dtype=param_dtype or torch.float32)

    def step():
        xi = x.detach().requires_grad_()
        if autocast is not None:
            with torch.autocast("cuda", dtype=autocast):
                e = net(xi, 500).sample
        else:
            e = net(xi, 500).sample
        (v,) = torch.autograd.grad(e, xi, grad_outputs=g.to(e.dtype))
        return v

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        step()
    e.record()
    cpu_issue = (time.perf_counter() - t0) / iters * 1e3
    torch.cuda.synchronize()
    print(f"{name:40s} gpu {s.elapsed_time(e) / iters:8.2f} ms/step   cpu-issue {cpu_issue:7.2f} ms   "
          f"mem {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB", flush=True)
    del net
    torch.cuda.empty_cache()


if __name__ == "__main__":
    run("fp32 NCHW", False, False)
    run("fp32 channels_last", True, False)
    run("fp32 channels_last + matmul tf32", True, True)
    run("bf16 autocast channels_last (info)", True, True, torch.bfloat16)
    run("bf16 autocast NCHW (info)", False, True, torch.bfloat16)
    run("bf16 weights NCHW (info)", False, True, param_dtype=torch.bfloat16)
    run("bf16 weights channels_last (info)", True, True, param_dtype=torch.bfloat16)
