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
NCU_TRAFFIC_FILE = os.path.join(ROOT, "profiles", "r02q_ncu_dram_bytes.json")   # written from the ncu --set full capture
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


def _ncu_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum of K1 + K2 per step from the committed `ncu --set full` capture at
    this exact shape (profiles/): (bytes, source) or (None, why)."""
    try:
        with open(NCU_TRAFFIC_FILE) as f:
            d = json.load(f)
        return int(d["k1_bytes"] + d["k2_bytes"]), d.get("source", NCU_TRAFFIC_FILE)
    except Exception as exc:
        return None, f"no capture on file ({type(exc).__name__})"


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
    # the tensor-core blur runs the pair with the bridge mean written by K1 (psx_dps_pre_mean / psx_dps_post_mean):
    # one wrapper for both launch modes
    real_pre_mean, real_post_mean = _native.dps_pre_mean, _native.dps_post_mean

    def timed_any(name, fn):
        graph_wrap, eager_wrap = timed_dev(name, fn), timed(name, fn)

        def wrap(*a, **k):
            return (graph_wrap if torch.cuda.is_current_stream_capturing() else eager_wrap)(*a, **k)
        return wrap

    dps_mod._native.dps_pre_mean = timed_any("k1", real_pre_mean)
    dps_mod._native.dps_post_mean = timed_any("k2", real_post_mean)

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
        # libpsx counts its own kernel launches: those recorded into the graph (= per replay) or of one eager step
        if use_graph:
            launches_per_step = run.graph_kernel_launches
        for k in range(W):
            if not use_graph and k == W - 1:
                l0 = _native.kernel_launches()
            run.step(k)
        if not use_graph:
            launches_per_step = _native.kernel_launches() - l0
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

        # ---------------- terminal exchange (untimed per step; reported once): final network pass + Tweedie into the
        # gather slot, then the all-gather / all-reduce
        counts = [L] * world
        gathered = torch.zeros((world * L, n), device=device)
        slot = gathered[rank * L:(rank + 1) * L]
        tot, tsq = torch.empty(n, device=device), torch.empty(n, device=device)
        barrier()
        g0, g1, g2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        g0.record()
        run.finalize(out=slot, total=tot, total_sq=tsq)
        g1.record()
        summary = combine_posterior(slot, gathered, tot, tsq, counts)
        g2.record()
        barrier()
        final_net_ms, final_collective_ms = g0.elapsed_time(g1), g1.elapsed_time(g2)
        assert summary.samples.shape[0] == world * L and torch.isfinite(summary.mean).all()
    finally:
        sampler.release()
        dps_mod._native.dps_pre, dps_mod._native.dps_post = real_pre, real_post
        dps_mod._native.dps_pre_dev, dps_mod._native.dps_post_dev = real_pre_dev, real_post_dev
        dps_mod._native.dps_pre_mean, dps_mod._native.dps_post_mean = real_pre_mean, real_post_mean

    t = torch.tensor([ms, ms_e2e, k1_ms, k2_ms], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e, k1_ms, k2_ms = (float(v) for v in t)
    step_s, e2e_s = ms / K / 1e3, ms_e2e / K / 1e3
    peak, peak_src = _peaks()
    alg_bytes = 40 * L * n                      # K1: 16 B/elem, K2: 24 B/elem
    fused_s = (k1_ms + k2_ms) / 1e3
    achieved = alg_bytes / fused_s / 1e9
    # ---------------- everything below is outside the timed regions; the UNet and its buffers are gone by now
    fused_mean = bool(getattr(run, "_fused_mean", False))
    del run, sampler, net, problem, summary, gathered
    torch.cuda.empty_cache()
    extra = {}
    if world == 1 and not args.no_extra:
        extra["by_config"] = roofline_by_config(device, peak)
        extra["aten_reference_step"] = aten_reference_step(device)
        extra["cfg1_gpu"] = config1_gpu(device)
    if world > 1:
        extra["multi_gpu_parity"] = multi_gpu_parity(device, rank, world)
    cpu = cpu_baseline(args) if (rank == 0 and world == 1 and not args.no_cpu_baseline) else None
    traffic, traffic_src = _ncu_traffic()

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
                   "parallelism": f"independent samples x{world}",
                   "terminal_final_network_pass_ms": final_net_ms, "terminal_gather_and_reduce_ms": final_collective_ms,
                   "launch": ("one CUDA-graph replay per timestep (scalars + timestep from device tables)"
                              if use_graph else "eager launches"),
                   **({"cuda_graph_error": graph_error} if graph_error else {}),
                   **({"cfg1_gpu": extra["cfg1_gpu"]} if "cfg1_gpu" in extra else {}),
                   **({"aten_reference_step": extra["aten_reference_step"]} if "aten_reference_step" in extra else {}),
                   **({"multi_gpu_parity": extra["multi_gpu_parity"]} if "multi_gpu_parity" in extra else {})},
        "e2e": {"value": world * L / (GUIDED_STEPS * e2e_s), "unit": "samples/s",
                "h2d_bytes_per_step": L * n * 4, "d2h_bytes_per_step": L * 4, "ms_per_step": ms_e2e / K},
        # libpsx's own count (psx_kernel_launches) of the kernels it launches per timestep, times the timed steps
        "gpu_launches": int(launches_per_step * K), "gpu_launches_per_step": int(launches_per_step),
        "abi_calls": int(abi_calls),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic if (L, n) == (16, 3 * 256 * 256) else None, "traffic_source": traffic_src,
                     "peak_source": peak_src,
                     "kernel": ("fused DPS step = K1 (blur_k1_tc<0, 1>: tcgen05 / TMEM, one launch; its spare CTAs on the "
                                "SMs the planes leave idle write the bridge mean) + K2 (k2_post_mean, on that mean)"
                                if fused_mean else
                                "fused DPS step = K1 (blur_k1_tc: tcgen05 / TMEM, one launch) + K2 (k2_post_v4)"),
                     "algorithmic_bytes_note": "40 B per element: x_t, eps, y in / cot out (K1), x_t, eps, cot, vjp, z "
                                               "in / x_next out (K2) -- the reference's data flow; the bridge-mean pair "
                                               "moves 4 of K2's bytes under K1 and is charged the same 40",
                     "algorithmic_bytes": alg_bytes, "k1_ms": k1_ms, "k2_ms": k2_ms,
                     "kernel_timing": ("external CUDA events inside the replayed graph, mean of K replays read "
                                       "one by one right after the timed region; last timed step: "
                                       f"k1 {k_last['k1']:.4f} ms, k2 {k_last['k2']:.4f} ms") if use_graph
                                      else "CUDA events around every ABI call inside the timed region",
                     "k2_alone_gbs": 24 * L * n / (k2_ms / 1e3) / 1e9,
                     "share_of_step": (k1_ms + k2_ms) / (ms / K),
                     **({"by_config": extra["by_config"]} if "by_config" in extra else {})},
        "clocks": clocks.summary(),
    }
    if cpu is not None:
        line["cpu_baseline"] = cpu
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


# =============================================================================== extra legs of the own arm
def _rotate_time(fn, nsets: int, passes: int) -> float:
    """ms per call: launches back to back over a ring of buffer sets whose footprint is >= 4x the 126 MB L2 (every launch
    reads cold inputs).  The ring pass is recorded ONCE into a CUDA graph and replayed, so that the host's launch path
    (Python -> ctypes -> cudaLaunchKernel, ~10 us per call: as long as a 12 us kernel) is not what is measured; one
    CUDA-event pair around each replay, median over the passes."""
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for i in range(nsets):
            fn(i)
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        for i in range(nsets):
            fn(i)
    ts = []
    for _ in range(passes):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        s.record()
        graph.replay()
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e) / nsets)
    del graph
    return statistics.median(ts)


def roofline_by_config(device, peak: float) -> list:
    """Kernel-only K1 / K2 of the other configurations through the C ABI (config 3: mask and 4x box at L = 64; the
    pointwise operators and the blur at both batch sizes): algorithmic bytes (16 / 24 B per element) over the measured
    launch time, as a fraction of the measured HBM peak."""
    from samplers_b200 import _native, operators as pops
    out = []
    gen = torch.Generator(device=device).manual_seed(0)
    makers = {"identity": lambda: pops.IdentityOperator(SHAPE),
              "mask70": lambda: pops.RandomInpaintingOperator(SHAPE, 0.7, flatten=False),
              "box4": lambda: pops.BoxDownsampleOperator(SHAPE, 4),
              "maskbox4": lambda: pops.MaskedBoxDownsampleOperator(SHAPE, 4, missing_fraction=0.7),
              "gblur61": lambda: pops.GaussianBlurOperator(SHAPE, 61, 3.0),
              "motion61_30deg": lambda: pops.MotionBlurOperator(SHAPE, kernel_size=61, angle_deg=30.0)}
    for kind, L in (("gblur61", 16), ("gblur61", 64), ("identity", 16), ("identity", 64), ("mask70", 64), ("box4", 64),
                    ("maskbox4", 64), ("motion61_30deg", 16)):
        op = makers[kind]().to(device)
        nat = op._native_cached(device)
        n = nat.n
        nsets = max(2, -(-4 * 126 * 2 ** 20 // (7 * L * n * 4)))
        S = []
        for _ in range(nsets):
            d = {k: torch.randn(L, n, device=device, generator=gen) for k in ("x", "eps", "v", "z")}
            d.update(cot=torch.empty(L, n, device=device), out=torch.empty(L, n, device=device),
                     part=torch.empty(L, nat.err_parts, device=device))
            wsb = nat.workspace_bytes(L)
            d["ws"] = torch.empty(wsb // 4, device=device) if wsb else None
            S.append(d)
        y = torch.randn(1, nat.n_y, device=device, generator=gen)

        def k1(i):
            d = S[i]
            _native.dps_pre(nat, d["x"], d["eps"], y, L, 0.8, 0.6, 400.0, d["cot"], d["part"], d["ws"])

        def k2(i):
            d = S[i]
            _native.dps_post(d["x"], d["eps"], d["cot"], d["v"], d["z"], d["part"], nat.err_parts, n, 0.8, 0.6, 0.99,
                             0.01, 0.05, 1.0, d["out"], None)

        for i in range(nsets):
            k1(i); k2(i)
        m1, m2 = _rotate_time(k1, nsets, 10), _rotate_time(k2, nsets, 10)
        classic = {}
        if nat.fuses_mean(L):
            # the pair the sampler runs for this operator at this batch: CTAs on the SMs K1 leaves idle write the bridge
            # mean + std z, K2 reads it instead of x_t, eps and z (bit-identical results)
            classic = {"k1_us_classic_pair": m1 * 1e3, "k2_us_classic_pair": m2 * 1e3, "pair": "K1 + bridge mean / K2 "
                       "on the mean (psx_dps_pre_mean / psx_dps_post_mean); *_classic_pair = psx_dps_pre / psx_dps_post"}
            for d in S:
                d["mean"] = torch.empty(L, n, device=device)

            def k1m(i):
                d = S[i]
                _native.dps_pre_mean(nat, d["x"], d["eps"], y, L, 0.8, 0.6, 400.0, 0.99, 0.01, d["cot"], d["part"],
                                     d["mean"], d["ws"], z=d["z"], std=0.05)

            def k2m(i):
                d = S[i]
                _native.dps_post_mean(d["mean"], d["cot"], d["v"], None, d["part"], nat.err_parts, n, 0.6, 0.0,
                                      1.0, d["out"], None)

            for i in range(nsets):
                k1m(i); k2m(i)
            m1, m2 = _rotate_time(k1m, nsets, 10), _rotate_time(k2m, nsets, 10)
        b1, b2 = 16 * L * n, 24 * L * n
        if kind.startswith("motion"):
            classic = {"k1_bound": "fp32 FMA, not HBM: 2 passes x 256 tap slots (163 non-zero taps) per element = 44.7 us "
                                   "at the measured 36 TFMA/s for L = 16 (conv2d_rowseg2, packed FFMA2)"}
        out.append({"operator": kind, "L": L, "k1_us": m1 * 1e3, "k2_us": m2 * 1e3, **classic,
                    "k1_frac": b1 / m1 / 1e6 / peak, "k2_frac": b2 / m2 / 1e6 / peak,
                    "fused_gbs": (b1 + b2) / (m1 + m2) / 1e6, "fused_frac": (b1 + b2) / (m1 + m2) / 1e6 / peak,
                    "timing": f"ring of {nsets} cold buffer sets, the ring pass replayed as one CUDA graph; the L "
                              "reconstructions share ONE observation (as in configs 2-3), which stays in L2: of the 16 "
                              "algorithmic B/elem of K1, 4 do not come from HBM -- fractions above 1 are that"})
        del S
        torch.cuda.empty_cache()
    return out


def aten_reference_step(device) -> dict:
    """What the two kernels replace, on the SAME GPU: the reference algorithm's timestep (the oracle's literal
    restatement of samplers/samplers/dps.py:96-122: Tweedie, residual, log-likelihood, autograd.grad, bridge update,
    guidance) issued as eager ATen launches, with a one-multiply stand-in network so that only the sampler arithmetic
    is timed -- against K1 + K2 on the same shapes (config 2: L = 16, 3 x 256 x 256, Gaussian blur 61).  Checker-side
    measurement: nothing of it is on the product path."""
    from oracle import dps as odps
    from oracle.operators import OracleGaussianBlur
    from oracle.schedule import ddpm_linear_alphas_cumprod, leading_timesteps_ascending, padded_clipped_acp
    from samplers_b200 import _native, operators as pops
    L = BATCH_PER_GPU
    acp = padded_clipped_acp(ddpm_linear_alphas_cumprod()).to(device)
    ts = leading_timesteps_ascending(SAMPLING_STEPS).tolist()
    op = OracleGaussianBlur(SHAPE, 61, 3.0)
    for name in ("taps_h", "taps_v"):
        if hasattr(op, name):
            setattr(op, name, getattr(op, name).to(device))
    g = torch.Generator(device=device).manual_seed(3)
    x = torch.randn(L, *SHAPE, device=device, generator=g)
    y = torch.randn(1, *SHAPE, device=device, generator=g)
    z = torch.randn(L, *SHAPE, device=device, generator=g)
    net = lambda xt, t: xt * 0.5  # noqa: E731
    sig = torch.tensor(SIGMA_Y, device=device)

    def step():
        return odps.dps_step_autograd(net, x, t=ts[500], t_prev=ts[499], s=ts[0], acp=acp, op=op, y=y,
                                      noise_kind="gaussian", noise_param=sig, gamma=1.0, eta=1.0, z=z)
    for _ in range(3):
        step()
    n_rep = 10
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    s.record()
    for _ in range(n_rep):
        step()
    e.record()
    torch.cuda.synchronize()
    aten_us = s.elapsed_time(e) / n_rep * 1e3
    # the same arithmetic through libpsx, eager, same tensors
    pop = pops.GaussianBlurOperator(SHAPE, 61, 3.0).to(device)
    nat = pop._native_cached(device)
    n = nat.n
    xf, ef, vf, zf = x.view(L, n), (x * 0.5).view(L, n), torch.randn(L, n, device=device, generator=g), z.view(L, n)
    cot, out, part = torch.empty(L, n, device=device), torch.empty(L, n, device=device), torch.empty(L, nat.err_parts, device=device)
    wsb = nat.workspace_bytes(L)
    ws = torch.empty(wsb // 4, device=device) if wsb else None

    def ours():
        _native.dps_pre(nat, xf, ef, y.view(1, n), L, 0.8, 0.6, 400.0, cot, part, ws)
        _native.dps_post(xf, ef, cot, vf, zf, part, nat.err_parts, n, 0.8, 0.6, 0.99, 0.01, 0.05, 1.0, out, None)
    for _ in range(3):
        ours()
    torch.cuda.synchronize()
    s.record()
    for _ in range(n_rep):
        ours()
    e.record()
    torch.cuda.synchronize()
    ours_us = s.elapsed_time(e) / n_rep * 1e3
    return {"aten_eager_step_us": aten_us, "libpsx_k1_k2_us": ours_us, "ratio": aten_us / ours_us,
            "what": "sampler arithmetic of one config-2 timestep (stand-in network): reference algorithm as eager ATen "
                    "ops vs K1 + K2, same GPU, warm L2, 10 repetitions"}


def config1_gpu(device) -> dict:
    """BASELINE config 1 exactly, whole `sampler(problem)` on the GPU (identity, sigma 0.05, 3 x 64 x 64, 50 steps,
    batch 1, ddpm-celebahq-256 UNet at random init): wall time after one warm-up call."""
    from samplers_b200.inverse_problem import InverseProblem
    from samplers_b200.networks import DDPMNetwork
    from samplers_b200.noise import GaussianNoise
    from samplers_b200.operators import IdentityOperator
    from samplers_b200.samplers import DPSSampler
    shape = (3, 64, 64)
    net = DDPMNetwork.from_config("google/ddpm-celebahq-256", seed=1234, device=device)
    x_true = (torch.rand(shape, generator=torch.Generator().manual_seed(0)) * 2 - 1).to(device)
    y = x_true + (torch.randn(shape, generator=torch.Generator().manual_seed(1)) * SIGMA_Y).to(device)
    prob = InverseProblem(operator=IdentityOperator(x_shape=shape).to(device), observation=y, noise=GaussianNoise(sigma=SIGMA_Y))
    res = {}
    for graph in (False, True):
        sampler = DPSSampler(net, cuda_graph=graph)
        walls = []
        for _ in range(2):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            out = sampler(prob, num_sampling_steps=50, num_reconstructions=1, gamma=SIGMA_Y ** 2, eta=1.0)
            torch.cuda.synchronize()
            walls.append(time.perf_counter() - t0)
        assert torch.isfinite(out).all()
        res["cuda_graph" if graph else "eager"] = {"wall_s": walls[-1], "samples_per_s": 1.0 / walls[-1]}
    del net
    torch.cuda.empty_cache()
    res["workload"] = "cfg1: DPS identity sigma 0.05, 3x64x64, 50 steps, batch 1, whole sampler(problem) wall, 2nd call"
    return res


def multi_gpu_parity(device, rank: int, world: int) -> dict:
    """Run under torchrun after the timed region (tiny networks, < 2 s): (1) DPS reconstructions sharded over the ranks
    (sample_posterior: gather + moment all-reduce) equal the same reconstructions computed on one rank; (2) PSLD with
    `process_group` (batch-global norms all-reduced, one scalar per norm) equals the full-batch run of one rank,
    sample for sample; (3) the same for ReSample.  Errors are the max over ranks."""
    import torch.distributed as dist
    from samplers_b200 import operators as P
    from samplers_b200.distributed import sample_posterior
    from samplers_b200.inverse_problem import InverseProblem
    from samplers_b200.networks import DDPMNetwork
    from samplers_b200.networks.sd15 import StableDiffusionCondition, StableDiffusionNetwork
    from samplers_b200.noise import GaussianNoise
    from samplers_b200.samplers import DPSSampler, PSLDSampler
    tf32 = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
    res = {}
    try:
        # ---- (1) DPS
        shape, per, steps = (3, 32, 32), 2, 8
        R = per * world
        net = DDPMNetwork.from_config("tiny", device=device)
        op = P.GaussianBlurOperator(shape, 9, 1.5).to(device)
        g = torch.Generator().manual_seed(0)
        x_true = (torch.rand(shape, generator=g) * 2 - 1).to(device)
        y = op.apply(x_true[None])[0] + 0.05 * torch.randn(shape, generator=g).to(device)
        prob = InverseProblem(operator=op, observation=y, noise=GaussianNoise(sigma=0.05))
        tape = [torch.randn(R, *shape, generator=torch.Generator().manual_seed(100 + i)) for i in range(steps)]

        def run_dps(lo, hi, sharded):
            it = iter(tape)
            s = DPSSampler(net)
            s.draw = lambda sh, dev, dtype: next(it)[lo:hi].to(dev)
            if sharded:
                return sample_posterior(s, prob, num_reconstructions=R, num_sampling_steps=steps, gamma=0.05)
            r = s.prepare(prob, steps, hi - lo, 0.05, 1.0, None)
            try:
                for k in range(r.num_steps):
                    r.step(k)
                return r.finalize().view(hi - lo, *shape)
            finally:
                s.release()
        summ = run_dps(rank * per, (rank + 1) * per, True)
        full = run_dps(0, R, False)
        scale = float(full.abs().max())
        errs = torch.tensor([float((summ.samples - full).abs().max()) / scale,
                             float((summ.mean - full.mean(0)).abs().max()) / scale,
                             float((summ.variance - full.var(0, unbiased=True)).abs().max()) / max(float(full.var(0).max()), 1e-12)],
                            device=device, dtype=torch.float64)
        dist.all_reduce(errs, op=dist.ReduceOp.MAX)
        res["dps_sharded_vs_one_rank"] = {"reconstructions": R, "samples_rel_err": float(errs[0]),
                                          "mean_rel_err": float(errs[1]), "variance_rel_err": float(errs[2])}
        del net
        # ---- (2) PSLD, global-batch semantics over the ranks
        lnet = StableDiffusionNetwork.from_config("sd15-tiny", device=device)
        xs = (3, 64, 64)
        opl = P.GaussianBlurOperator(xs, 9, 1.5).to(device)
        xt = (torch.rand(xs, generator=torch.Generator().manual_seed(5)) * 2 - 1).to(device)
        yl = opl.apply(xt[None])[0] + 0.05 * torch.randn(xs, generator=torch.Generator().manual_seed(6)).to(device)
        probl = InverseProblem(operator=opl, observation=yl, noise=GaussianNoise(sigma=0.05))
        cond = StableDiffusionCondition(guidance_scale=1.0, prompt_embeds=torch.zeros(1, 7, 32))
        psteps = 6
        zshape = tuple(lnet.get_latent_shape(xs))
        ltape = [torch.randn(world, *zshape, generator=torch.Generator().manual_seed(200 + i)) for i in range(psteps)]

        def run_psld(lo, hi, group):
            it = iter(ltape)
            s = PSLDSampler(lnet, process_group=group)
            s.draw = lambda sh, dev, dtype: next(it)[lo:hi].to(dev)
            return s(probl, num_sampling_steps=psteps, num_reconstructions=hi - lo, condition=cond)
        mine = run_psld(rank, rank + 1, dist.group.WORLD).reshape(1, *xs)
        fullb = run_psld(0, world, None).reshape(world, *xs)
        e = torch.tensor([float((mine[0] - fullb[rank]).abs().max() / fullb.abs().max())], device=device, dtype=torch.float64)
        dist.all_reduce(e, op=dist.ReduceOp.MAX)
        res["psld_process_group_vs_full_batch"] = {"batch": world, "rel_err": float(e[0])}
        # ---- (3) ReSample, the same way (MSE / norm sums all-reduced per conditioning step and optimiser iteration)
        from samplers_b200.samplers import ReSampleSampler
        rkw = dict(num_sampling_steps=8, sigma_scale=40.0, max_optimization_iters=6, eta=1.0, inter_timesteps=2,
                   time_travel_interval=2, stage_splits=3)

        def run_resample(lo, hi, group):
            count = [0]

            def draw(sh, dev, dtype):   # draw i of the full batch, this rank's rows (every draw is batch-leading)
                count[0] += 1
                full = torch.randn((world,) + tuple(sh[1:]), generator=torch.Generator().manual_seed(300 + count[0]))
                return full[lo:hi].to(dev)
            s = ReSampleSampler(lnet, process_group=group)
            s.draw = draw
            return s(probl, num_reconstructions=hi - lo, condition=cond, **rkw)
        mine = run_resample(rank, rank + 1, dist.group.WORLD).reshape(1, *xs)
        fullb = run_resample(0, world, None).reshape(world, *xs)
        e = torch.tensor([float((mine[0] - fullb[rank]).abs().max() / fullb.abs().max())], device=device, dtype=torch.float64)
        dist.all_reduce(e, op=dist.ReduceOp.MAX)
        res["resample_process_group_vs_full_batch"] = {"batch": world, "rel_err": float(e[0])}
    except Exception as exc:   # reported, never silent
        res["error"] = f"{type(exc).__name__}: {exc}"[:300]
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = tf32
    return res


# =============================================================================== CPU arm: the oracle port
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


def _cpu_config1() -> dict:
    """BASELINE config 1 exactly on the host cores: the oracle port's whole sampler (identity, sigma 0.05, 3 x 64 x 64,
    50 steps, batch 1, the same random-init ddpm-celebahq-256 UNet), wall of the second call."""
    from oracle import dps as odps
    from oracle.operators import OracleIdentity
    from oracle.schedule import ddpm_linear_alphas_cumprod, leading_timesteps_ascending, padded_clipped_acp
    from samplers_b200.networks.unet2d import CELEBAHQ_256, UNet2DModel
    shape = (3, 64, 64)
    torch.manual_seed(1234)
    unet = UNet2DModel(**CELEBAHQ_256).eval().requires_grad_(False)
    net = lambda x, t: unet(x, t).sample  # noqa: E731
    acp = padded_clipped_acp(ddpm_linear_alphas_cumprod())
    ts = leading_timesteps_ascending(50).tolist()
    x_true = torch.rand(shape, generator=torch.Generator().manual_seed(0)) * 2 - 1
    y = (x_true + torch.randn(shape, generator=torch.Generator().manual_seed(1)) * SIGMA_Y)[None]
    walls = []
    for _ in range(2):
        gz = torch.Generator().manual_seed(2)
        t0 = time.perf_counter()
        out = odps.dps_sample(net, acp=acp, timesteps=ts, op=OracleIdentity(shape), y=y, noise_kind="gaussian",
                              noise_param=torch.tensor(SIGMA_Y), gamma=SIGMA_Y ** 2, eta=1.0, leading=1,
                              draw=lambda sh: torch.randn(sh, generator=gz))
        walls.append(time.perf_counter() - t0)
    assert torch.isfinite(out).all()
    return {"wall_s": walls[-1], "samples_per_s": 1.0 / walls[-1],
            "workload": "cfg1: DPS identity sigma 0.05, 3x64x64, 50 steps, batch 1, whole sampler wall, 2nd call"}


def _cpu_math_only() -> list:
    """The sampler arithmetic alone on the host cores (one-multiply stand-in network) at the config-2 / config-3 tensor
    shapes: ms per timestep and effective GB/s = 40 L C H W / time -- the CPU counterpart of the fused-step GB/s."""
    from oracle import dps as odps
    from oracle.operators import OracleBoxDownsample, OracleGaussianBlur
    from oracle.schedule import ddpm_linear_alphas_cumprod, leading_timesteps_ascending, padded_clipped_acp
    acp = padded_clipped_acp(ddpm_linear_alphas_cumprod())
    ts = leading_timesteps_ascending(SAMPLING_STEPS).tolist()
    net = lambda x, t: x * 0.5  # noqa: E731
    res = []
    for name, L, op in (("cfg2 gblur61", 16, OracleGaussianBlur(SHAPE, 61, 3.0)), ("cfg3 box4", 64, OracleBoxDownsample(SHAPE, 4))):
        g = torch.Generator().manual_seed(4)
        x = torch.randn(L, *SHAPE, generator=g)
        y = op.apply(torch.rand(1, *SHAPE, generator=g) * 2 - 1)
        ts_ = []
        for k in range(3):
            z = torch.randn(L, *SHAPE, generator=g)
            t0 = time.perf_counter()
            odps.dps_step_autograd(net, x, t=ts[500], t_prev=ts[499], s=ts[0], acp=acp, op=op, y=y, noise_kind="gaussian",
                                   noise_param=torch.tensor(SIGMA_Y), gamma=1.0, eta=1.0, z=z)
            ts_.append(time.perf_counter() - t0)
        sec = min(ts_[1:])
        res.append({"workload": name, "L": L, "ms_per_step": sec * 1e3, "gbs": 40 * L * 3 * 256 * 256 / sec / 1e9})
    return res


def _use_all_host_threads() -> int:
    """torchrun exports OMP_NUM_THREADS=1 to every rank; the CPU arm is meant to use all the host threads it can."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    torch.set_num_threads(max(1, n))
    return torch.get_num_threads()


def _host_mem_gib() -> float:
    try:
        import psutil
        return psutil.virtual_memory().available / 2 ** 30
    except Exception:
        return 0.0


def cpu_baseline(args) -> dict:
    """Bounded sample (~30 s of host work) reported beside the GPU line: one timestep of the timed workload at batch 1,
    config 1 end to end, and the sampler arithmetic alone at the config-2 / config-3 shapes."""
    _use_all_host_threads()
    times = _cpu_port_step_time(1, steps=1, warmup=1)
    sec = statistics.mean(times)
    return {"value": 1 / (GUIDED_STEPS * sec), "unit": "samples/s", "cores": torch.get_num_threads(),
            "host_cpus": os.cpu_count(), "kind": "port",
            "sample": f"1 timed DPS timestep (after 1 warm-up) of the same workload at batch 1 instead of 16, "
                      f"oracle/dps.py dps_step_autograd on CPU fp32; {sec:.2f} s/step "
                      "(--impl reference runs the full batch)",
            "cfg1_whole_sampler": _cpu_config1(), "sampler_math_only": _cpu_math_only()}


def run_reference(args):
    """The reference algorithm on the host cores (oracle port; the reference itself is Python and does not travel to the
    GPU box), SAME configuration as the own arm: batch 16, the driver's --steps / --warmup.  If the host cannot fit
    that in memory or in --ref-budget-s, the batch per step is reduced -- and the line says so (`same_config`)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = _use_all_host_threads()
    K, W = args.steps, args.warmup
    t_start = time.perf_counter()
    probe = statistics.mean(_cpu_port_step_time(1, steps=1, warmup=1))   # s per sample-step, calibration
    mem = _host_mem_gib()
    batch = BATCH_PER_GPU if args.cpu_batch <= 0 else args.cpu_batch
    why = None
    while batch > 1 and (probe * batch * (K + W) > args.ref_budget_s or (mem and 4.0 * batch > 0.6 * mem)):
        batch //= 2
        why = (f"host: {probe:.2f} s per sample-step, {mem:.0f} GiB free; batch {BATCH_PER_GPU} x {K + W} steps does not "
               f"fit {args.ref_budget_s:.0f} s / the memory")
    times = _cpu_port_step_time(batch, steps=K, warmup=W)
    sec = statistics.mean(times)
    value = batch / (GUIDED_STEPS * sec)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    cfg1, math_only = _cpu_config1(), _cpu_math_only()
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "samples/s", "n_gpus": world,
        "steps": K, "warmup": W, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "batch_per_step": batch, "same_config": batch == BATCH_PER_GPU,
                   **({"reduced_because": why} if why else {}), "shape": list(SHAPE),
                   "sampling_steps": SAMPLING_STEPS, "guided_steps_per_sample": GUIDED_STEPS,
                   "timed_steps": len(times), "warmup_steps": W, "wall_s": time.perf_counter() - t_start,
                   "cfg1_whole_sampler": cfg1, "sampler_math_only": math_only,
                   "note": "reference algorithm (oracle port of samplers/samplers/dps.py, literal autograd step) with the "
                           "same random-init ddpm-celebahq-256 UNet on the host CPU, one process, all host threads"},
        "cpu_baseline": {"value": value, "unit": "samples/s", "cores": threads, "host_cpus": os.cpu_count(),
                         "kind": "port", "sample": f"{len(times)} timed DPS timesteps at batch {batch}, {sec:.2f} s/step"},
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
    ap.add_argument("--cpu-batch", type=int, default=0)          # reference arm: 0 = the own arm's batch (16)
    ap.add_argument("--ref-budget-s", type=float, default=420.0)  # reference arm: host time the timed steps may take
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true")           # skip by_config / ATen step / config 1 legs
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "own":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_own(args)


if __name__ == "__main__":
    main()
