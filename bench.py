#!/usr/bin/env python
"""bench.py -- DPS posterior samples/s at 256^2 + fused-step HBM GB/s (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (config.workload): BASELINE.json configs[1] -- DPSSampler, Gaussian-blur deblurring
(61x61, sigma 3.0), ddpm-celebahq-256 UNet (random init), 1000 sampling steps, batch 16 per GPU,
3x256x256, fp32, synthetic observation.  One "step" = one guided DPS timestep over the batch
(UNet forward, K1, UNet VJP, K2).  A posterior sample costs 998 such steps (dps.py:91), so

    value [samples/s] = n_gpus * 16 / (998 * seconds_per_step)

Independent samples are sharded across ranks (weak scaling): no collective inside the loop; the
terminal all-gather + moment all-reduce is run once after the timed region and reported in config.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SHAPE = (3, 256, 256)
BATCH_PER_GPU = 16
SAMPLING_STEPS = 1000
GUIDED_STEPS = SAMPLING_STEPS - 2        # range(len(ts)-1, 1, -1)
SIGMA_Y = 0.05
# dram__bytes_read.sum + dram__bytes_write.sum per step from the committed ncu --set full capture of the four
# launches at this exact shape (L = 16): 25.19 + 13.41 + 12.61 (K1) + 62.93 + 1.25 (K2) MB
NCU_DRAM_BYTES_PER_STEP = int((25.19 + 13.41 + 12.60 + 62.93 + 2.06) * 1e6)  # rows + cols + rows_il + K2 read, K2 write
METRIC = "dps_posterior_samples_per_s_256"
WORKLOAD = "cfg2: DPS gaussian-blur 61x61 sigma3, ddpm-celebahq-256 UNet, 1000 steps, batch 16/GPU, 3x256x256"


def _peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (profiling recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.rows, self.proc, self.index = [], None, index

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None
        return self

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *exc):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self) -> dict:
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for nm, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                continue
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                "samples": len(sm)}


def build_problem(device, batch: int):
    """Synthetic config-2 problem: x_true ~ U[-1,1] (seed 0), y = A x + N(0, 0.05^2) (seed 1)."""
    from samplers_b200.inverse_problem import InverseProblem
    from samplers_b200.noise import GaussianNoise
    from samplers_b200.operators import GaussianBlurOperator
    op = GaussianBlurOperator(SHAPE, 61, 3.0).to(device)
    x_true = (torch.rand(SHAPE, generator=torch.Generator().manual_seed(0)) * 2 - 1).to(device)
    clean = op(x_true)
    noise = GaussianNoise(sigma=SIGMA_Y)
    y = clean + (torch.randn(clean.shape, generator=torch.Generator().manual_seed(1)) * SIGMA_Y).to(device)
    return InverseProblem(operator=op, observation=y, noise=noise)


# =============================================================================== own arm (CUDA)
def run_own(args):
    import torch.distributed as dist
    from samplers_b200 import _native
    from samplers_b200.networks import DDPMNetwork
    from samplers_b200.samplers import DPSSampler
    from samplers_b200.distributed import combine_posterior

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback")
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    _native.load()
    torch.backends.cudnn.benchmark = bool(args.cudnn_benchmark)

    L = BATCH_PER_GPU
    net = DDPMNetwork.from_config("google/ddpm-celebahq-256", seed=1234, device=device,
                                  channels_last=bool(args.channels_last))
    problem = build_problem(device, L)
    sampler = DPSSampler(net)
    torch.manual_seed(2 + rank)                 # the N(0,1) draws come from torch's CUDA generator on every path
    run = sampler.prepare(problem, num_sampling_steps=SAMPLING_STEPS, num_reconstructions=L, gamma=1.0, eta=1.0)
    n = run.n
    K, W = args.steps, args.warmup
    use_graph, graph_error = bool(args.cuda_graph), None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # per-kernel CUDA events on the launching (torch current) stream
    ev = {name: [] for name in ("k1", "k2")}
    real_pre, real_post = _native.dps_pre, _native.dps_post
    timing = {"on": False}

    def timed(name, fn):
        def wrap(*a, **k):
            if not timing["on"]:
                return fn(*a, **k)
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record(); fn(*a, **k); e.record()
            ev[name].append((s, e))
        return wrap

    import samplers_b200.samplers.dps as dps_mod
    dps_mod._native.dps_pre = timed("k1", real_pre)
    dps_mod._native.dps_post = timed("k2", real_post)

    # CUDA-graph replay of the timestep: K1 / K2 are bracketed by EXTERNAL events recorded inside the graph
    gev = {}
    real_pre_dev, real_post_dev = _native.dps_pre_dev, _native.dps_post_dev

    def timed_dev(name, fn):
        def wrap(*a, **k):
            if not torch.cuda.is_current_stream_capturing():
                return fn(*a, **k)
            s, e = (torch.cuda.Event(enable_timing=True, external=True) for _ in range(2))
            s.record(); fn(*a, **k); e.record()
            gev[name] = (s, e)
        return wrap

    dps_mod._native.dps_pre_dev = timed_dev("k1", real_pre_dev)
    dps_mod._native.dps_post_dev = timed_dev("k2", real_post_dev)

    try:
        if use_graph:
            try:
                run.capture(draw_in_graph=False)   # noise buffer filled per step: on device (value) / from host (e2e)
            except Exception as exc:               # reported in the JSON line, never silent
                graph_error = f"{type(exc).__name__}: {exc}"[:200]
                use_graph = False
                torch.cuda.synchronize()
                sampler.release()
                run = sampler.prepare(problem, num_sampling_steps=SAMPLING_STEPS, num_reconstructions=L,
                                      gamma=1.0, eta=1.0)
        for k in range(W):
            run.step(k)
        # ---------------- device-resident timing (value)
        barrier()
        timing["on"] = True
        launches0 = _native.launch_count
        with ClockSampler(local) as clocks:
            t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0.record()
            for k in range(K):
                run.step((W + k) % run.num_steps)
            t1.record()
            barrier()
        timing["on"] = False
        ms = t0.elapsed_time(t1)
        abi_calls = _native.launch_count - launches0
        if use_graph:
            # the in-graph events now hold the LAST timed step; K more replays, read after each, give the mean
            k_last = {name: s.elapsed_time(e) for name, (s, e) in gev.items()}
            samples = {"k1": [], "k2": []}
            for k in range(K):
                run.step((W + K + k) % run.num_steps)
                torch.cuda.synchronize()
                for name, (s, e) in gev.items():
                    samples[name].append(s.elapsed_time(e))
            k1_ms, k2_ms = statistics.mean(samples["k1"]), statistics.mean(samples["k2"])
            abi_calls = 2 * K                      # the two ABI calls are nodes of the replayed graph
        else:
            k_last = None
            k1_ms = statistics.mean(s.elapsed_time(e) for s, e in ev["k1"])
            k2_ms = statistics.mean(s.elapsed_time(e) for s, e in ev["k2"])

        # ---------------- end-to-end through the public step API with HOST buffers
        z_host = [torch.randn(run.view.flat_shape, generator=torch.Generator().manual_seed(100 + i)).pin_memory()
                  for i in range(2)]
        err_host = torch.empty(L, dtype=torch.float32).pin_memory()
        # graph replay reads its noise from run.z: copy straight into it
        z_dev = run.z.view(run.view.flat_shape) if use_graph else torch.empty(run.view.flat_shape, device=device)
        for k in range(2):  # warm the copy path
            z_dev.copy_(z_host[k % 2], non_blocking=True); run.step((W + K + k) % run.num_steps, z=z_dev)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for k in range(K):
            z_dev.copy_(z_host[k % 2], non_blocking=True)          # H2D: this step's injected noise
            run.step((W + K + 2 + k) % run.num_steps, z=z_dev)
            err_host.copy_(run.err, non_blocking=True)             # D2H: per-sample |y - A x0|
        e1.record()
        barrier()
        ms_e2e = e0.elapsed_time(e1)

        # ---------------- terminal exchange (untimed per step; reported once)
        counts = [L] * world
        gathered = torch.zeros((world * L, n), device=device)
        slot = gathered[rank * L:(rank + 1) * L]
        tot, tsq = torch.empty(n, device=device), torch.empty(n, device=device)
        barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        run.finalize(out=slot, total=tot, total_sq=tsq)
        summary = combine_posterior(slot, gathered, tot, tsq, counts)
        g1.record()
        barrier()
        final_ms = g0.elapsed_time(g1)
        assert summary.samples.shape[0] == world * L and torch.isfinite(summary.mean).all()
    finally:
        sampler.release()
        dps_mod._native.dps_pre, dps_mod._native.dps_post = real_pre, real_post
        dps_mod._native.dps_pre_dev, dps_mod._native.dps_post_dev = real_pre_dev, real_post_dev

    t = torch.tensor([ms, ms_e2e, k1_ms, k2_ms], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e, k1_ms, k2_ms = (float(v) for v in t)
    step_s, e2e_s = ms / K / 1e3, ms_e2e / K / 1e3
    peak, peak_src = _peaks()
    alg_bytes = 40 * L * n                      # K1: 16 B/elem, K2: 24 B/elem
    fused_s = (k1_ms + k2_ms) / 1e3
    achieved = alg_bytes / fused_s / 1e9
    cpu = cpu_baseline(args) if (rank == 0 and world == 1 and not args.no_cpu_baseline) else None

    line = {
        "metric": METRIC, "value": world * L / (GUIDED_STEPS * step_s), "unit": "samples/s",
        "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms / K, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "batch_per_gpu": L, "global_batch": world * L, "shape": list(SHAPE),
                   "sampling_steps": SAMPLING_STEPS, "guided_steps_per_sample": GUIDED_STEPS,
                   "step": "one guided DPS timestep over the batch: UNet fwd + K1 + UNet VJP + K2 (+ randn)",
                   "l2": "inputs larger than L2: the UNet pass between consecutive K1/K2 touches GBs of activations",
                   "network": "ddpm-celebahq-256 UNet2D, random init, torch (cuDNN TF32 conv defaults"
                              + (", channels_last" if args.channels_last else "") + ")",
                   "parallelism": f"independent samples x{world}", "terminal_gather_ms": final_ms,
                   "launch": ("one CUDA-graph replay per timestep (scalars + timestep from device tables)"
                              if use_graph else "eager launches"),
                   **({"cuda_graph_error": graph_error} if graph_error else {})},
        "e2e": {"value": world * L / (GUIDED_STEPS * e2e_s), "unit": "samples/s",
                "h2d_bytes_per_step": L * n * 4, "d2h_bytes_per_step": L * 4, "ms_per_step": ms_e2e / K},
        # K1 = 3 kernels per sample group; inside a captured graph the blur K1 runs as two groups (launch_pre_sepblur)
        "gpu_launches": int((_native.KERNELS_PER_CALL["pre_sepblur"] * (2 if use_graph and L % 2 == 0 and L >= 4 else 1)
                             + _native.KERNELS_PER_CALL["post"]) * K),
        "abi_calls": int(abi_calls),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": NCU_DRAM_BYTES_PER_STEP if (L, n) == (16, 3 * 256 * 256) else None,
                     "traffic_source": "profiles/r01_ncu_full_summary.csv (ncu --set full, cold cache: "
                                       "dram read+write of the 3 K1 launches + K2)",
                     "peak_source": peak_src, "kernel": "fused DPS step = K1 (3 kernels per sample group) + K2",
                     "algorithmic_bytes": alg_bytes, "k1_ms": k1_ms, "k2_ms": k2_ms,
                     "kernel_timing": ("external CUDA events inside the replayed graph, mean of K replays read "
                                       "one by one right after the timed region; last timed step: "
                                       f"k1 {k_last['k1']:.4f} ms, k2 {k_last['k2']:.4f} ms") if use_graph
                                      else "CUDA events around every ABI call inside the timed region",
                     "k2_alone_gbs": 24 * L * n / (k2_ms / 1e3) / 1e9,
                     "share_of_step": (k1_ms + k2_ms) / (ms / K)},
        "clocks": clocks.summary(),
    }
    if cpu is not None:
        line["cpu_baseline"] = cpu
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


# =============================================================================== CPU oracle port
def _cpu_port_step_time(batch: int, steps: int, warmup: int):
    """Times the oracle's literal DPS step (autograd form) on the host cores, same network + operator."""
    from oracle import dps as odps
    from oracle.operators import OracleGaussianBlur
    from oracle.schedule import ddpm_linear_alphas_cumprod, leading_timesteps_ascending, padded_clipped_acp
    from samplers_b200.networks.unet2d import CELEBAHQ_256, UNet2DModel

    torch.manual_seed(1234)
    unet = UNet2DModel(**CELEBAHQ_256).eval().requires_grad_(False)
    net = lambda x, t: unet(x, t).sample  # noqa: E731
    acp = padded_clipped_acp(ddpm_linear_alphas_cumprod())
    ts = leading_timesteps_ascending(SAMPLING_STEPS).tolist()
    op = OracleGaussianBlur(SHAPE, 61, 3.0)
    x_true = torch.rand(SHAPE, generator=torch.Generator().manual_seed(0)) * 2 - 1
    y = op.apply(x_true[None]) + torch.randn(1, *SHAPE, generator=torch.Generator().manual_seed(1)) * SIGMA_Y
    g = torch.Generator().manual_seed(2)
    x = torch.randn(batch, *SHAPE, generator=g)
    times = []
    for k in range(warmup + steps):
        i = len(ts) - 1 - k
        z = torch.randn(batch, *SHAPE, generator=g)
        t0 = time.perf_counter()
        out = odps.dps_step_autograd(net, x, t=ts[i], t_prev=ts[i - 1], s=ts[0], acp=acp, op=op, y=y,
                                     noise_kind="gaussian", noise_param=torch.tensor(SIGMA_Y), gamma=1.0,
                                     eta=1.0, z=z)
        dt = time.perf_counter() - t0
        if k >= warmup:
            times.append(dt)
        x = out["x_next"]
    return times


def _use_all_host_threads() -> int:
    """torchrun exports OMP_NUM_THREADS=1 to every rank; the CPU arm is meant to use all the host threads it can."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    torch.set_num_threads(max(1, n))
    return torch.get_num_threads()


def cpu_baseline(args) -> dict:
    _use_all_host_threads()
    batch = args.cpu_batch
    times = _cpu_port_step_time(batch, steps=1, warmup=1)
    sec = statistics.mean(times)
    return {"value": batch / (GUIDED_STEPS * sec), "unit": "samples/s", "cores": torch.get_num_threads(),
            "host_cpus": os.cpu_count(), "kind": "port",
            "sample": f"1 timed DPS timestep (after 1 warm-up) of the same workload at batch {batch} instead of 16, "
                      f"oracle/dps.py dps_step_autograd on CPU fp32; {sec:.2f} s/step"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    _use_all_host_threads()
    batch = args.cpu_batch
    K, W = args.steps, args.warmup
    K_eff, W_eff = min(K, args.ref_max_steps), min(W, 1)
    t0 = time.perf_counter()
    times = _cpu_port_step_time(batch, steps=K_eff, warmup=W_eff)
    sec = statistics.mean(times)
    value = batch / (GUIDED_STEPS * sec)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "samples/s", "n_gpus": world,
        "steps": K, "warmup": W, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "batch_per_step": batch, "shape": list(SHAPE),
                   "sampling_steps": SAMPLING_STEPS, "guided_steps_per_sample": GUIDED_STEPS,
                   "timed_steps": K_eff, "warmup_steps": W_eff, "wall_s": time.perf_counter() - t0,
                   "note": "reference algorithm (oracle port of samplers/samplers/dps.py) on the host CPU; each step is a "
                           "bounded sample of the workload (reduced batch); samples/s scales per sample"},
        "cpu_baseline": {"value": value, "unit": "samples/s", "cores": torch.get_num_threads(),
                         "host_cpus": os.cpu_count(), "kind": "port",
                         "sample": f"{K_eff} timed DPS timesteps at batch {batch}, {sec:.2f} s/step"},
        "e2e": {"value": value, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="own", choices=["own", "reference"])
    ap.add_argument("--channels-last", type=int, default=int(os.environ.get("PSX_CHANNELS_LAST", "0")))  # NCHW measured 1.36x faster (tools/unet_bench.py)
    ap.add_argument("--cudnn-benchmark", type=int, default=1)
    ap.add_argument("--cuda-graph", type=int, default=int(os.environ.get("PSX_CUDA_GRAPH", "1")))  # 1.03x at config 2 (tools/graph_bench.py)
    ap.add_argument("--cpu-batch", type=int, default=1)
    ap.add_argument("--ref-max-steps", type=int, default=4)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "own":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_own(args)


if __name__ == "__main__":
    main()
