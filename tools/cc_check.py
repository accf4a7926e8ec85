#!/usr/bin/env python
"""Development check of the CUDA-core blur K1 (PSX_NO_TC=1) at large batches: per-sample error against fp64 for
L x obs_repeat x PSX_SPLIT; prints which samples are off."""
from __future__ import annotations

import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))

from samplers_b200 import _native, operators as pops  # noqa: E402
from tc_check import ref64, run  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    gen = torch.Generator(device=dev).manual_seed(0)
    op = pops.GaussianBlurOperator((3, 256, 256), 61, 3.0).to(dev)
    nat = op._native_cached(dev)
    os.environ["PSX_NO_TC"] = "1"
    bad_any = False
    for L, rep in [(32, 1), (32, 32), (48, 48), (64, 1), (64, 64), (64, 8)]:
        n = nat.n
        x = torch.randn(L, n, device=dev, generator=gen)
        eps = torch.randn(L, n, device=dev, generator=gen)
        y = torch.randn(L // rep, nat.n_y, device=dev, generator=gen)
        rc, re = ref64(op, x, eps, y, 0.9, 0.43, 25.0, rep)
        for split, extra in (("1", None), ("2", None), ("1", "PSX_NO_FAST16"), ("1", "PSX_NO_PIPE")):
            os.environ["PSX_SPLIT"] = split
            os.environ.pop("PSX_NO_FAST16", None)
            os.environ.pop("PSX_NO_PIPE", None)
            if extra:
                os.environ[extra] = "1"
            _native.reload_env()
            cot, err = run(nat, x, eps, y, L, 0.9, 0.43, 25.0, rep)
            cot2, err2 = run(nat, x, eps, y, L, 0.9, 0.43, 25.0, rep)
            same = bool((cot == cot2).all())
            per = ((cot.double() - rc).norm(dim=1) / rc.norm(dim=1)).cpu()
            bad = [int(i) for i in torch.nonzero(per > 1e-5).flatten()]
            bad_any |= bool(bad)
            print(json.dumps({"L": L, "obs_repeat": rep, "split": split, "extra": extra, "max_rel": float(per.max()),
                              "err_rel": float(((err - re).abs() / re).max()), "bad_samples": bad, "repeatable": same}), flush=True)
    return 1 if bad_any else 0


if __name__ == "__main__":
    sys.exit(main())
