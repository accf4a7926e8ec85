#!/usr/bin/env python
"""Summaries of ncu output for profiles/.

  python tools/ncu_summarize.py launches gpurun_out/launches.csv  > profiles/rNN_ncu_bench_step_launches.csv
      input: `ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file ... python bench.py ...`;
      output: one guided DPS timestep (the launches between the last two k2_post kernels) aggregated by kernel name.
  python tools/ncu_summarize.py full raw.csv > profiles/rNN_ncu_full_summary.csv
      input: `ncu -i X.ncu-rep --page raw --csv`; output: the columns DESIGN.md / bench.py quote, one row per launch.
"""
from __future__ import annotations

import csv
import io
import json
import sys
from collections import OrderedDict


def read_csv(path):
    lines = [ln for ln in open(path, newline="") if ln.startswith('"')]
    return list(csv.reader(io.StringIO("".join(lines))))


def launches(path):
    rows = read_csv(path)
    hdr, rows = rows[0], rows[1:]
    ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    seq = []
    for r in rows:
        v = float(r[iv].replace(",", ""))
        v = v / 1e3 if r[iu] in ("ns", "nsecond") else v * (1e3 if r[iu] in ("ms", "msecond") else 1.0)
        seq.append((r[ik], v))
    k2 = [i for i, (k, _) in enumerate(seq) if "k2_post" in k]
    if len(k2) < 2:
        raise SystemExit("fewer than two k2_post launches in the capture")
    step = seq[k2[-2] + 1:k2[-1] + 1]
    agg = OrderedDict()
    for k, v in step:
        a = agg.setdefault(k[:80], [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    out = csv.writer(sys.stdout)
    print(f"# one guided DPS timestep = the {len(step)} launches between the last two k2_post kernels, {tot / 1e3:.2f} ms of "
          "kernel time, aggregated by kernel name (ncu times are cold-cache and serialised: compare shares)")
    out.writerow(["kernel", "launches", "total_us", "share"])
    for k, (n, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.writerow([k, n, f"{v:.2f}", f"{v / tot:.5f}"])
    psx = {k: a for k, a in agg.items() if "psx::" in k}
    print("# libpsx: " + json.dumps({k[:60]: {"launches": n, "us": round(v, 2)} for k, (n, v) in psx.items()}) +
          f" = {sum(a[1] for a in psx.values()) / tot:.5%} of the step")


FULL_COLS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
             "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
             "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
             "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_active",
             "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
             "launch__block_size", "launch__cluster_size", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
             "smsp__inst_executed.sum", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
             "lts__throughput.avg.pct_of_peak_sustained_elapsed"]


def full(path):
    rows = read_csv(path)
    hdr, units, rows = rows[0], rows[1], rows[2:]
    ik = hdr.index("Kernel Name")
    cols = [(c, hdr.index(c)) for c in FULL_COLS if c in hdr]
    out = csv.writer(sys.stdout)
    out.writerow(["kernel"] + [f"{c} [{units[i]}]" for c, i in cols])
    for r in rows:
        out.writerow([r[ik][:60]] + [r[i] for _, i in cols])


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
