"""Randomised geometry sweep (fixed seeds): for every operator kind and ~40 random (C, H, W, L) the stand-alone
forward / adjoint match the oracle and the fused K1 equals w * A^T (y - A x0) / sa assembled from the stand-alone
kernels -- the generic kernels, the pipelined kernels and the strip kernels are all reached by some shape."""
import random

import pytest
import torch

from oracle import operators as oops
from tests._golden import rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _cases(kind, count, seed):
    rng = random.Random(seed)
    out = []
    for _ in range(count):
        c = rng.choice([1, 2, 3, 4])
        if kind == "box":
            f = rng.choice([2, 3, 4, 8])
            h, w = f * rng.randint(1, 12), f * rng.randint(1, 12)
            extra = f
        elif kind == "blur":
            # a mix: strip-kernel geometries (multiples of 32 / 16), pipelined ones (W % 8 == 0) and fully ragged
            mode = rng.choice(["strip", "pipe", "ragged"])
            if mode == "strip":
                h, w = 16 * rng.randint(1, 20), 32 * rng.randint(1, 10)
            elif mode == "pipe":
                h, w = rng.randint(5, 90), 8 * rng.randint(1, 20)
            else:
                h, w = rng.randint(3, 70), rng.randint(3, 70)
            extra = (rng.choice([3, 5, 9, 13, 21, 31]), rng.uniform(0.6, 3.0))
        elif kind == "sep":   # arbitrary (asymmetric, different per direction) separable taps, incl. the K = 64 kernels
            mode = rng.choice(["strip", "pipe", "ragged", "big"])
            if mode == "strip":
                h, w = 16 * rng.randint(1, 16), 32 * rng.randint(1, 8)
            elif mode == "pipe":
                h, w = rng.randint(5, 90), 8 * rng.randint(1, 20)
            elif mode == "big":
                h, w = 32 * rng.randint(9, 16), 32 * rng.randint(9, 16)
            else:
                h, w = rng.randint(3, 70), rng.randint(3, 70)
            extra = (rng.choice([1, 3, 7, 15, 33, 61]), rng.choice([1, 5, 15, 39, 61]), rng.randint(0, 10 ** 6))
        elif kind == "motion":
            h, w = rng.randint(3, 80), rng.randint(3, 140)
            extra = (rng.choice([5, 9, 15]), rng.uniform(0.0, 180.0))
        else:
            h, w = rng.randint(1, 40), rng.randint(1, 40)
            extra = rng.random()
        out.append((c, h, w, rng.randint(1, 6), extra))
    return out


def _build(kind, shape, extra):
    from samplers_b200 import operators as P
    from samplers_b200.operators.blur import motion_line_kernel
    if kind == "identity":
        return P.IdentityOperator(shape), oops.OracleIdentity(shape), None
    if kind == "mask":
        op = P.RandomInpaintingOperator(shape, 0.3 + 0.5 * extra, seed=int(extra * 1000), flatten=False)
        keep = (~op.mask).float()

        class Dense(oops.OracleOperator):
            def apply(self, x):
                return x * keep.to(x.device)

            adjoint = apply

        return op, Dense(), keep
    if kind == "box":
        return P.BoxDownsampleOperator(shape, extra), oops.OracleBoxDownsample(shape, extra), None
    if kind == "blur":
        return P.GaussianBlurOperator(shape, extra[0], extra[1]), oops.OracleGaussianBlur(shape, extra[0], extra[1]), None
    if kind == "sep":
        g = torch.Generator().manual_seed(extra[2])
        th, tv = torch.rand(extra[0], generator=g) + 0.05, torch.rand(extra[1], generator=g) + 0.05
        th, tv = th / th.sum(), tv / tv.sum()
        return P.SeparableBlurOperator(shape, th, tv), oops.OracleSeparableBlur(shape, th, tv), None
    k2d = motion_line_kernel(extra[0], extra[1])
    return P.MotionBlurOperator(shape, kernel=k2d), oops.OracleConv2dBlur(shape, k2d), None


def test_blur_heights_between_128_and_256_regression():
    """conv_cols_pipe used to assume (H / 8) * 16 tasks == rounds * 256 exactly: heights 144 .. 240 lost their lower rows
    in the stand-alone A / A^T (found by the sweep below)."""
    from samplers_b200 import operators as P
    torch.backends.cudnn.allow_tf32 = False
    for h in (144, 192, 240):
        shape = (1, h, 160)
        op = P.GaussianBlurOperator(shape, 21, 2.0).to(DEV)
        ora = oops.OracleGaussianBlur(shape, 21, 2.0)
        ora.taps_h, ora.taps_v = ora.taps_h.to(DEV), ora.taps_v.to(DEV)
        x = torch.randn(2, *shape, device=DEV, generator=torch.Generator(device=DEV).manual_seed(h))
        assert rel_err(op.apply(x).cpu(), ora.apply(x).cpu()) < 3e-6
        assert rel_err(op.apply_transpose(x).cpu(), ora.adjoint(x).cpu()) < 3e-6


@pytest.mark.parametrize("kind,env", [("identity", None), ("mask", None), ("box", None), ("blur", None),
                                      ("blur", "PSX_NO_FAST16"), ("blur", "PSX_NO_PIPE"), ("blur", "PSX_SPLIT"),
                                      ("sep", None), ("sep", "PSX_NO_FAST16"), ("motion", None)])
def test_random_geometries(monkeypatch, kind, env):
    from samplers_b200 import _native
    torch.backends.cudnn.allow_tf32 = False
    for k in ("PSX_NO_FAST16", "PSX_NO_PIPE", "PSX_SPLIT", "PSX_FUSED", "PSX_PSF_FORM"):
        monkeypatch.delenv(k, raising=False)
    if env:
        monkeypatch.setenv(env, "2")     # any value switches the NO_* paths; PSX_SPLIT=2 forces two sample groups
    gen = torch.Generator(device=DEV).manual_seed(sum(map(ord, kind)))
    for c, h, w, L, extra in _cases(kind, 40 if kind != "sep" else 30, seed=len(kind)):
        shape = (c, h, w)
        op, ora, keep = _build(kind, shape, extra)
        op = op.to(DEV)
        for name in ("taps_h", "taps_v", "kernel2d"):
            if hasattr(ora, name):
                setattr(ora, name, getattr(ora, name).to(DEV))
        nat = op._native_cached(torch.device(DEV))
        tag = f"{kind} {shape} L={L} {extra}"
        x = torch.randn(L, *shape, device=DEV, generator=gen)
        ax = op.apply(x)
        assert rel_err(ax.cpu(), ora.apply(x).reshape(ax.shape).cpu()) < 3e-6, tag
        yv = torch.randn(ax.shape, device=DEV, generator=gen)
        aty = op.apply_transpose(yv)
        assert rel_err(aty.cpu(), ora.adjoint(yv.reshape(L, *ora.y_shape) if hasattr(ora, "y_shape") else yv)
                       .reshape(aty.shape).cpu()) < 3e-6, tag
        eps = torch.randn(L, nat.n, device=DEV, generator=gen)
        y = torch.randn(1, nat.n_y, device=DEV, generator=gen)
        if keep is not None:
            y = y * keep.reshape(1, -1).to(DEV)
        sa, s1, wgt = 0.8, 0.6, 400.0
        xf = x.reshape(L, -1).contiguous()
        cot, part = torch.empty(L, nat.n, device=DEV), torch.empty(L, nat.err_parts, device=DEV)
        wsb = nat.workspace_bytes(L)
        ws = torch.empty(wsb // 4, device=DEV) if wsb else None
        _native.dps_pre(nat, xf, eps, y, L, sa, s1, wgt, cot, part, ws)
        x0 = torch.empty_like(xf)
        _native.tweedie(xf, eps, sa, s1, x0)
        r = y - nat.apply(x0)
        assert rel_err(cot.cpu(), (nat.adjoint(r.contiguous()) * wgt / sa).cpu()) < 3e-6, tag
        assert rel_err(part.sum(1).cpu(), r.double().square().sum(1).float().cpu()) < 1e-5, tag
