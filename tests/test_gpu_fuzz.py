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
        if kind in ("box", "maskbox"):
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
    if kind == "maskflat":   # the reference's working inpainting form: gathered (L, m) observations
        mask = torch.rand(shape, generator=torch.Generator().manual_seed(int(extra * 1000))) < 0.3 + 0.5 * extra
        mask.view(-1)[0] = False                                   # at least one kept pixel
        return P.InpaintingOperator(shape, mask, flatten=True), oops.OracleMaskGather(shape, mask), None
    if kind == "box":
        return P.BoxDownsampleOperator(shape, extra), oops.OracleBoxDownsample(shape, extra), None
    if kind == "maskbox":   # config 3 as one operator: mask on the coarse grid after the box average
        op = P.MaskedBoxDownsampleOperator(shape, extra, missing_fraction=0.6, seed=shape[1] * 131 + shape[2])
        return op, oops.OracleMaskedBox(shape, extra, (~op.mask).float()), None
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


@pytest.mark.parametrize("kind,env", [("identity", None), ("mask", None), ("box", None), ("maskbox", None), ("blur", None),
                                      ("blur", "PSX_NO_FAST16"), ("blur", "PSX_NO_PIPE"), ("blur", "PSX_SPLIT"),
                                      ("sep", None), ("sep", "PSX_NO_FAST16"), ("motion", None)])
def test_random_geometries(psx_env, kind, env):
    from samplers_b200 import _native
    torch.backends.cudnn.allow_tf32 = False
    psx_env(PSX_NO_FAST16=None, PSX_NO_PIPE=None, PSX_SPLIT=None, PSX_FUSED=None, PSX_PSF_FORM=None, PSX_NO_TC=None)
    if env:
        # any value switches the NO_* paths; PSX_SPLIT=2 forces two sample groups.  These are switches of the CUDA-core
        # strip kernels: keep the tensor-core kernel (which 256 x 256 planes would take first) out of the way.
        psx_env(**{env: "2", "PSX_NO_TC": "1"})
    gen = torch.Generator(device=DEV).manual_seed(sum(map(ord, kind)))
    for c, h, w, L, extra in _cases(kind, 40 if kind != "sep" else 30, seed=len(kind)):
        shape = (c, h, w)
        op, ora, keep = _build(kind, shape, extra)
        op = op.to(DEV)
        for name in ("taps_h", "taps_v", "kernel2d", "kept"):
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


def test_random_sizes_elementwise_family():
    """K2 (tensor noise / none / fixed scale / Philox-vs-tensor), the final Tweedie with moments, bridge update,
    lincomb3, eps-DDIM, stochastic resample, gather / scatter at ~60 random (L, n) incl. n % 4 != 0, against the
    oracle on the CPU (the schedule scalars must come from CPU torch ops, as in the reference's CPU path and the
    host planner: sqrt of a 0-dim tensor differs between CPU and CUDA torch by an ulp for ~1 % of the timesteps)."""
    from oracle import dps as odps
    from oracle import resample as ors
    from oracle.schedule import ddpm_linear_alphas_cumprod, padded_clipped_acp
    from samplers_b200 import _native
    from samplers_b200.samplers.resample import ddim_eps_scalars, resample_scalars
    acp = padded_clipped_acp(ddpm_linear_alphas_cumprod())
    rng = random.Random(7)
    gen = torch.Generator(device=DEV).manual_seed(7)
    for _ in range(60):
        L = rng.randint(1, 7)
        n = rng.choice([1, 2, 3, 5, 17, 64, 100, 1001, 4096, 12289, 3 * 64 * 64, rng.randint(1, 30000)])
        t = rng.randint(2, 999)
        tp = rng.randint(1, t - 1)
        a_t = acp[t]
        sa, s1 = float(a_t ** 0.5), float((1 - a_t) ** 0.5)
        x, e, d, v, z = (torch.randn(L, n, device=DEV, generator=gen) for _ in range(5))
        xc, ec, dc, vc, zc = (q.cpu() for q in (x, e, d, v, z))
        tag = f"L={L} n={n} t={t}"
        # ---- K2
        parts = rng.choice([1, 3, 64])
        part = torch.rand(L, parts, device=DEV, generator=gen) * n
        c_ell, c_s, std, gamma = 0.9 + 0.1 * rng.random(), 0.05 * rng.random(), 0.3 * rng.random(), 0.5 + rng.random()
        out, err = torch.empty(L, n, device=DEV), torch.empty(L, device=DEV)
        _native.dps_post(x, e, d, v, z, part, parts, n, sa, s1, c_ell, c_s, std, gamma, out, err)
        ref, _ = odps.k2_reference(xc, ec, dc, vc, zc, part.sum(1).cpu(), acp_t=a_t, c_ell=torch.tensor(c_ell),
                                   c_s=torch.tensor(c_s), std=torch.tensor(std), gamma=gamma)
        assert rel_err(out.cpu(), ref) < 2e-6, tag
        assert rel_err(err.cpu(), part.sum(1).sqrt().cpu()) < 1e-6, tag
        _native.dps_post(x, e, d, v, None, part, parts, n, sa, s1, c_ell, c_s, 0.0, gamma, out, None)
        ref0, _ = odps.k2_reference(xc, ec, dc, vc, None, part.sum(1).cpu(), acp_t=a_t, c_ell=torch.tensor(c_ell),
                                    c_s=torch.tensor(c_s), std=torch.tensor(0.0), gamma=gamma)
        assert rel_err(out.cpu(), ref0) < 2e-6, tag
        zp = torch.empty(L, n, device=DEV)
        _native.philox_normal(zp, 5, t)
        a, b = torch.empty(L, n, device=DEV), torch.empty(L, n, device=DEV)
        _native.dps_post(x, e, d, v, zp, None, 0, n, sa, s1, c_ell, c_s, std, gamma, a, None)      # fixed scale
        _native.dps_post_philox(x, e, d, v, None, 0, n, sa, s1, c_ell, c_s, std, gamma, 5, t, b, None)
        assert torch.equal(a, b), tag
        # ---- final Tweedie + moments
        x0, tot, tsq = torch.empty(L, n, device=DEV), torch.empty(n, device=DEV), torch.empty(n, device=DEV)
        _native.tweedie(x, e, sa, s1, x0, tot, tsq)
        want = odps.tweedie_x0(xc, ec, a_t).to(DEV)
        assert torch.equal(x0, want), tag
        assert rel_err(tot.cpu(), want.double().sum(0).float().cpu()) < 1e-6, tag
        assert rel_err(tsq.cpu(), want.double().square().sum(0).float().cpu()) < 1e-6, tag
        # ---- bridge update / lincomb3
        _native.bridge_update(x, e, z, d, sa, s1, c_ell, c_s, std, -1.0, out)
        assert rel_err(out.cpu(), (c_ell * x + c_s * want + std * z - d).cpu()) < 2e-6, tag
        _native.lincomb3(x, 0.3, e, -1.7, d, 2.0, out)
        assert rel_err(out.cpu(), (0.3 * x - 1.7 * e + 2.0 * d).cpu()) < 2e-6, tag
        # ---- eps-DDIM / stochastic resample
        prev, pseudo = torch.empty_like(x), torch.empty_like(x)
        _native.ddim_eps_step(x, e, z, ddim_eps_scalars(acp, t, tp, 0.7), prev, None, pseudo)
        rp, _, rps = ors.ddim_eps_step(xc, ec, acp, t, tp, 0.7, zc)
        assert rel_err(prev.cpu(), rp) < 2e-6 and rel_err(pseudo.cpu(), rps) < 2e-6, tag
        sig = ors.compute_sigma(40.0, acp[t], acp[tp])
        _native.stochastic_resample(x, e, z, *resample_scalars(acp, t, tp, 40.0), out)
        assert rel_err(out.cpu(), ors.stochastic_resample(xc, ec, acp[tp], sig, zc)) < 2e-6, tag
        # ---- gather / scatter
        m = rng.randint(0, n)
        idx = torch.sort(torch.randperm(n, generator=torch.Generator().manual_seed(n))[:m]).values.to(DEV)
        if m:
            gth = _native.gather(x, idx, False, n)
            assert torch.equal(gth, x.index_select(1, idx)), tag
            sct = _native.gather(gth, idx, True, n)
            chk = torch.zeros(L, n, device=DEV)
            chk[:, idx] = gth
            assert torch.equal(sct, chk), tag


def test_random_sampler_configurations_one_step():
    """DPSRun.step against the oracle's literal autograd step (CPU) for ~36 random (operator, batch shape,
    reconstructions, noise model, eta, gamma, timestep) -- the observation tiling, the likelihood weight and the bridge
    coefficients all vary."""
    from oracle import dps as odps
    from oracle.schedule import ddpm_linear_alphas_cumprod, leading_timesteps_ascending, padded_clipped_acp
    from oracle.tiny_net import TinyEpsNet
    from samplers_b200 import operators as P
    from samplers_b200.inverse_problem import InverseProblem
    from samplers_b200.noise import GaussianNoise, PoissonNoise
    from samplers_b200.samplers import DPSSampler
    from tests.test_gpu_fullsize import _network
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    acp = padded_clipped_acp(ddpm_linear_alphas_cumprod())
    rng = random.Random(11)
    for case in range(36):
        steps = rng.choice([10, 20, 50])
        ts = leading_timesteps_ascending(steps)
        shape = (3, 8 * rng.randint(1, 5), 8 * rng.randint(1, 5))
        kind = rng.choice(["identity", "mask", "maskflat", "box", "blur", "motion"])
        extra = {"identity": None, "mask": rng.random(), "maskflat": rng.random(), "box": rng.choice([2, 4]),
                 "blur": (rng.choice([5, 9]), 1.2), "motion": (5, rng.uniform(0, 180))}[kind]
        op, ora, keep = _build(kind, shape, extra)
        batch = rng.choice([(), (2,), (3,)])
        R = rng.choice([1, 2, 3])
        nb = 1
        for b in batch:
            nb *= b
        L = nb * R
        noise_kind, param = rng.choice([("gaussian", 0.05), ("gaussian", 0.2), ("poisson", 4.0)])
        eta, gamma = rng.choice([0.0, 0.5, 1.0]), rng.choice([0.05, 0.3, 1.0])
        g = torch.Generator().manual_seed(case)
        y = torch.randn(*batch, *ora.y_shape if hasattr(ora, "y_shape") else shape, generator=g)
        if keep is not None:
            y = y * keep
        net = _network(acp, ts)
        cpu_core = TinyEpsNet(channels=3)                      # same deterministic weights as net.core
        x_init = torch.randn(L, *shape, generator=g)
        z = torch.randn(L, *shape, generator=g)
        noise = GaussianNoise(sigma=param) if noise_kind == "gaussian" else PoissonNoise(rate=param)
        prob = InverseProblem(operator=op.to(DEV), observation=y.to(DEV), noise=noise)
        s = DPSSampler(net)
        s.draw = lambda shape_, device, dtype: x_init.to(device).view(shape_)
        run = s.prepare(prob, num_sampling_steps=steps, num_reconstructions=R, gamma=gamma, eta=eta)
        try:
            k = rng.randint(0, run.num_steps - 1)
            sc = run.plan[k]
            run.step(k, z=z.to(DEV))
            # the oracle sees the observation tiled over reconstructions (BatchView.repeat_observation)
            y_l = y.reshape(nb, *y.shape[len(batch):]).repeat_interleave(R, dim=0)
            ref = odps.dps_step_autograd(lambda xx, tt: cpu_core(xx, int(tt)), x_init, t=sc.t, t_prev=sc.t_prev,
                                         s=run.timesteps[0], acp=acp, op=ora, y=y_l, noise_kind=noise_kind,
                                         noise_param=torch.tensor(param), gamma=gamma, eta=eta, z=z)
            tag = f"case {case}: {kind} {shape} batch={batch} R={R} {noise_kind} eta={eta} gamma={gamma} k={k}"
            assert rel_err(run.x.view(L, *shape).cpu(), ref["x_next"]) < 2e-5, tag
            assert rel_err(run.err.cpu(), ref["err"]) < 1e-5, tag
        finally:
            s.release()


def test_random_geometries_latent_sampler_autograd_nodes():
    """The PSLD data-term node (lik, x_eff) and ReSample's residual node (norm / mse), forward and backward, against
    the oracle's torch expressions through autograd, for ~30 random operators / shapes / batch sizes."""
    from samplers_b200.samplers.psld import _PsldDataTerm
    from samplers_b200.samplers.resample import _ResidualTerm
    torch.backends.cudnn.allow_tf32 = False
    rng = random.Random(5)
    gen = torch.Generator(device=DEV).manual_seed(5)
    for case in range(30):
        kind = rng.choice(["identity", "mask", "box", "blur", "motion"])
        c, h, w, L, extra = _cases(kind, 1, seed=100 + case)[0]
        shape = (c, h, w)
        op, ora, keep = _build(kind, shape, extra)
        op = op.to(DEV)
        for name in ("taps_h", "taps_v", "kernel2d"):
            if hasattr(ora, name):
                setattr(ora, name, getattr(ora, name).to(DEV))
        nat = op._native_cached(torch.device(DEV))
        tag = f"case {case}: {kind} {shape} L={L}"
        x0 = torch.randn(L, *shape, device=DEV, generator=gen)
        y = torch.randn(1, nat.n_y, device=DEV, generator=gen)
        if keep is not None:
            y = y * keep.reshape(1, -1).to(DEV)
        cx = torch.randn(L, nat.n, device=DEV, generator=gen)
        wsb = nat.workspace_bytes(L)
        ws = torch.empty(wsb // 4, device=DEV) if wsb else None
        zeros_y = torch.zeros(1, nat.n_y, device=DEV)
        # oracle expressions (psld.py:129-136, resample_kernels.py:27,48)
        xo = x0.clone().requires_grad_()
        hx = ora.apply(xo).reshape(L, -1)
        r_o = y - hx
        lik_o = torch.norm(r_o)
        aty = ora.adjoint(y.expand(L, -1).reshape(L, *getattr(ora, "y_shape", shape))).reshape(L, -1)
        xeff_o = aty + xo.reshape(L, -1) - ora.adjoint(hx.reshape(L, *getattr(ora, "y_shape", shape))).reshape(L, -1)
        (g_o,) = torch.autograd.grad(0.3 * lik_o + (xeff_o * cx).sum() + 2.0 * r_o.square().mean(), xo)
        # product nodes
        xd = x0.reshape(L, -1).clone().requires_grad_()
        lik, xeff = _PsldDataTerm.apply(xd, nat, y, L, ws, zeros_y)
        mse = _ResidualTerm.apply(xd, nat, y, L, ws, "mse")
        nrm = _ResidualTerm.apply(xd, nat, y, L, ws, "norm")
        (g_d,) = torch.autograd.grad(0.3 * lik + (xeff * cx).sum() + 2.0 * mse, xd)
        scale = max(1.0, float(lik_o.detach()))
        assert abs(float(lik.detach()) - float(lik_o.detach())) < 1e-5 * scale, tag
        assert abs(float(nrm.detach()) - float(lik_o.detach())) < 1e-5 * scale, tag
        assert abs(float(mse.detach()) - float(r_o.detach().square().mean())) < 1e-5 * max(1.0, float(mse.detach())), tag
        assert rel_err(xeff.detach().cpu(), xeff_o.detach().cpu()) < 5e-6, tag
        assert rel_err(g_d.cpu(), g_o.reshape(L, -1).cpu()) < 2e-5, tag


def test_random_geometries_bf16_state_is_the_rounded_fp32_path():
    """K_bf16(v) == bf16_rn(K_fp32(float(v))) for ~40 random operators / shapes (identity, mask, 4x box, separable blur
    incl. planes up to 512 and forced two-group schedules), K1 and K2."""
    import os
    from samplers_b200 import _native, operators as P
    BF = torch.bfloat16
    rng = random.Random(21)
    gen = torch.Generator(device=DEV).manual_seed(21)
    for case in range(40):
        kind = rng.choice(["identity", "mask", "box", "blur", "blur"])
        c, L = rng.choice([1, 2, 3]), rng.randint(1, 6)
        if kind == "box":
            h, w = 4 * rng.randint(1, 20), 4 * rng.randint(1, 20)
            op = P.BoxDownsampleOperator((c, h, w), 4)
        elif kind == "blur":
            h, w = 16 * rng.randint(1, 32), 32 * rng.randint(1, 16)
            if (h > 256 or w > 256) and h % 32:
                h += 16
            op = P.GaussianBlurOperator((c, h, w), *rng.choice([(61, 3.0), (9, 1.5), (13, 1.5)]))
        else:
            h, w = rng.randint(1, 50), rng.randint(1, 50)
            op = P.IdentityOperator((c, h, w)) if kind == "identity" else P.RandomInpaintingOperator((c, h, w), 0.6, seed=case, flatten=False)
        op = op.to(DEV)
        nat = op._native_cached(torch.device(DEV))
        n = nat.n
        tag = f"case {case}: {kind} {(c, h, w)} L={L}"
        os.environ.pop("PSX_SPLIT", None)
        os.environ["PSX_NO_TC"] = "1"   # the rounding identity is between the bf16 and the fp32 strip kernels
        if kind == "blur" and L % 2 == 0 and L >= 4 and rng.random() < 0.5:
            os.environ["PSX_SPLIT"] = "2"
        _native.reload_env()
        try:
            x, e, v, z = (torch.randn(L, n, device=DEV, generator=gen).to(BF) for _ in range(4))
            y = torch.randn(1, nat.n_y, device=DEV, generator=gen)
            if kind == "mask":
                y = y * (~op.mask).float().reshape(1, -1).to(DEV)
            sa, s1, wgt = 0.8366600275039673, 0.547722578048706, 400.0
            wsb = nat.workspace_bytes(L)
            ws = torch.empty(wsb // 4, device=DEV) if wsb else None
            c32, p32 = torch.empty(L, n, device=DEV), torch.empty(L, nat.err_parts, device=DEV)
            c16, p16 = torch.empty(L, n, device=DEV, dtype=BF), torch.empty(L, nat.err_parts, device=DEV)
            _native.dps_pre(nat, x.float(), e.float(), y, L, sa, s1, wgt, c32, p32, ws)
            _native.dps_pre_bf16(nat, x, e, y, L, sa, s1, wgt, c16, p16, ws=ws)
            assert torch.equal(c16, c32.to(BF)), tag
            assert torch.allclose(p16.sum(1), p32.sum(1), rtol=1e-6), tag
            o32, o16 = torch.empty(L, n, device=DEV), torch.empty(L, n, device=DEV, dtype=BF)
            _native.dps_post(x.float(), e.float(), c16.float(), v.float(), z.float(), p16, nat.err_parts, n, sa, s1, 0.97,
                             0.02, 0.1, 1.3, o32, None)
            _native.dps_post_bf16(x, e, c16, v, z, p16, nat.err_parts, n, sa, s1, 0.97, 0.02, 0.1, 1.3, o16, None)
            assert torch.equal(o16, o32.to(BF)), tag
        finally:
            os.environ.pop("PSX_SPLIT", None)
            os.environ.pop("PSX_NO_TC", None)
            _native.reload_env()


def test_random_pgdm_runs():
    """PGDMSampler (K1 with weight 2c, fixed-scale K2) against the oracle's whole PGDM run (CPU) for ~12 random
    (operator with a pseudo-inverse, shape, reconstructions, guidance weight, eta); eager and graph."""
    from oracle import pgdm as opg
    from oracle.schedule import ddpm_linear_alphas_cumprod, leading_timesteps_ascending, padded_clipped_acp
    from oracle.tiny_net import TinyEpsNet
    from samplers_b200.inverse_problem import InverseProblem
    from samplers_b200.noise import GaussianNoise
    from samplers_b200.samplers import PGDMSampler
    from tests.test_gpu_fullsize import _network
    torch.backends.cudnn.allow_tf32 = False
    acp = padded_clipped_acp(ddpm_linear_alphas_cumprod())
    rng = random.Random(13)
    for case in range(12):
        steps = rng.choice([6, 8, 12])
        ts = leading_timesteps_ascending(steps)
        kind = rng.choice(["identity", "maskflat", "box"])
        shape = (3, 8 * rng.randint(1, 4), 8 * rng.randint(1, 4))
        extra = {"identity": None, "maskflat": rng.random(), "box": rng.choice([2, 4])}[kind]
        op, ora, _ = _build(kind, shape, extra)
        R = rng.choice([1, 2, 3])
        gw, eta = rng.choice([0.2, 0.5, 1.0]), rng.choice([0.0, 1.0])
        g = torch.Generator().manual_seed(200 + case)
        y = torch.randn(*ora.y_shape, generator=g)
        draws = [torch.randn(R, *shape, generator=g) for _ in range(steps)]
        net = _network(acp, ts)
        cpu_core = TinyEpsNet(channels=3)
        it = iter(draws)
        ref = opg.pgdm_sample(lambda xx, tt: cpu_core(xx, int(tt)), acp=acp, timesteps=ts.tolist(), op=ora,
                              y_flat=y.reshape(1, *ora.y_shape).expand(R, *ora.y_shape), leading=R, guidance_weight=gw,
                              eta=eta, draw=lambda sh: next(it))
        prob = InverseProblem(operator=op.to(DEV), observation=y.to(DEV), noise=GaussianNoise(sigma=0.05))
        it2 = iter(draws)
        s = PGDMSampler(net)
        s.draw = lambda sh, device, dtype: next(it2).to(device)
        out = s(prob, num_sampling_steps=steps, num_reconstructions=R, guidance_weight=gw, eta=eta,
                keep_reconstruction_dim=True)
        tag = f"case {case}: {kind} {shape} R={R} gw={gw} eta={eta} steps={steps}"
        assert out.shape == (R, *shape), tag
        assert rel_err(out.cpu(), ref) < 1e-4, tag


def test_random_psld_runs():
    """PSLDSampler against the oracle's whole PSLD run (CPU) for ~8 random (operator, shape, batch, reconstructions,
    omega, gamma, eta) with the tiny deterministic latent network."""
    from oracle import psld as ops_
    from oracle.schedule import ddpm_linear_alphas_cumprod, leading_timesteps_ascending, padded_clipped_acp
    from oracle.tiny_latent_net import TinyLatentCore
    from samplers_b200.inverse_problem import InverseProblem
    from samplers_b200.networks.base import LatentEpsilonNetwork
    from samplers_b200.noise import GaussianNoise
    from samplers_b200.samplers import PSLDSampler
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    acp = padded_clipped_acp(ddpm_linear_alphas_cumprod())

    class Net(LatentEpsilonNetwork):
        def __init__(self, ts):
            super().__init__(alphas_cumprod=acp)
            self.core, self._ts = TinyLatentCore(channels=3), ts

        def forward(self, x, t):
            return self.core.eps(x, int(t))

        @classmethod
        def from_pretrained(cls, *a, **k):
            raise NotImplementedError

        def set_sampling_parameters(self, num_sampling_steps, batch_size=1, num_reconstructions=1):
            self._batch_size, self._num_sampling_steps = batch_size, num_sampling_steps
            self._num_reconstructions = num_reconstructions
            self.register_buffer("timesteps", self._ts.to(self.alphas_cumprod.device))

        @property
        def is_condition_initialized(self):
            return True

        def get_latent_shape(self, x_shape):
            return (4, x_shape[1] // 2, x_shape[2] // 2)

        def _decode(self, z, *, differentiable=False):
            return self.core.decode(z)

        def _encode(self, x, *, differentiable=False):
            return self.core.encode(x)

    rng = random.Random(17)
    for case in range(8):
        steps = rng.choice([5, 6, 8])
        ts = leading_timesteps_ascending(steps)
        kind = rng.choice(["identity", "box", "blur"])
        shape = (3, 8 * rng.randint(1, 3), 8 * rng.randint(1, 3))
        extra = {"identity": None, "box": 2, "blur": (rng.choice([5, 9]), 1.2)}[kind]
        op, ora, _ = _build(kind, shape, extra)
        batch = rng.choice([(), (2,)])
        R = rng.choice([1, 2])
        nb = 2 if batch else 1
        L = nb * R
        omega, gamma, eta = rng.choice([0.1, 0.3]), rng.choice([0.5, 1.0]), rng.choice([0.0, 0.5, 1.0])
        g = torch.Generator().manual_seed(300 + case)
        y = torch.randn(*batch, *ora.y_shape, generator=g)
        lat = (4, shape[1] // 2, shape[2] // 2)
        draws = [torch.randn(L, *lat, generator=g) for _ in range(steps)]
        core = TinyLatentCore(channels=3)
        it = iter(draws)
        y_flat = y.reshape(nb, *ora.y_shape).repeat_interleave(R, dim=0)
        ref = ops_.psld_sample(core.eps, core.decode, core.encode, acp=acp, timesteps=ts.tolist(), op=ora, y_flat=y_flat,
                               latent_shape=lat, leading=L, omega=omega, gamma=gamma, eta=eta, draw=lambda sh: next(it))
        net = Net(ts).to(DEV)
        prob = InverseProblem(operator=op.to(DEV), observation=y.to(DEV), noise=GaussianNoise(sigma=0.05))
        it2 = iter(draws)
        s = PSLDSampler(net)
        s.draw = lambda sh, device, dtype: next(it2).to(device)
        out = s(prob, num_sampling_steps=steps, num_reconstructions=R, gamma=gamma, omega=omega, eta=eta)
        tag = f"case {case}: {kind} {shape} batch={batch} R={R} omega={omega} gamma={gamma} eta={eta}"
        assert out.shape == (*batch, R, *shape), tag
        assert rel_err(out.reshape(L, *shape).cpu(), ref) < 2e-4, tag
