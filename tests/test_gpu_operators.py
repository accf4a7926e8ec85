"""CUDA operator kernels (through the C ABI) vs the CPU oracle operators, plus
size-independent properties at BASELINE.json's full sizes (L=16, 3x256x256)."""
import pytest
import torch

from oracle import operators as oops
from tests._golden import rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _pairs(shape, seed=3):
    from samplers_b200 import operators as pops
    g = torch.Generator().manual_seed(seed)
    mask = torch.rand(shape, generator=g) < 0.7
    th, tv = torch.rand(7, generator=g), torch.rand(5, generator=g)
    line, walk = oops.motion_line_kernel(11, 30.0), oops.motion_walk_kernel(13, 0.5, seed=1)
    c, h, w = shape
    out = {
        "identity": (pops.IdentityOperator(shape), oops.OracleIdentity(shape)),
        "identity_flat": (pops.IdentityOperator(shape, flatten=True), oops.OracleIdentity(shape, flatten=True)),
        "mask_gather": (pops.InpaintingOperator(shape, mask), oops.OracleMaskGather(shape, mask)),
        "gblur9": (pops.GaussianBlurOperator(shape, 9, 1.5), oops.OracleGaussianBlur(shape, 9, 1.5)),
        "gblur61": (pops.GaussianBlurOperator(shape), oops.OracleGaussianBlur(shape, 61, 3.0)),
        "sep_asym": (pops.SeparableBlurOperator(shape, th, tv), oops.OracleSeparableBlur(shape, th, tv)),
        "motion_line": (pops.MotionBlurOperator(shape, kernel=line), oops.OracleConv2dBlur(shape, line)),
        "motion_walk": (pops.MotionBlurOperator(shape, kernel=walk), oops.OracleConv2dBlur(shape, walk)),
    }
    for f in (2, 3, 4, 8):
        if h % f == 0 and w % f == 0:
            out[f"box{f}"] = (pops.BoxDownsampleOperator(shape, f), oops.OracleBoxDownsample(shape, f))
            cm = torch.rand(c, h // f, w // f, generator=g) < 0.7   # True = missing coarse pixel
            out[f"maskbox{f}"] = (pops.MaskedBoxDownsampleOperator(shape, f, mask=cm),
                                  oops.OracleMaskedBox(shape, f, (~cm).float()))
    return out


SHAPES = [(3, 24, 48), (1, 17, 23), (3, 64, 64), (2, 40, 264)]  # ragged sizes on purpose


@pytest.mark.parametrize("shape", SHAPES)
def test_apply_and_adjoint_match_oracle(shape):
    for name, (op, ora) in _pairs(shape).items():
        g = torch.Generator().manual_seed(11)
        x = torch.randn(3, *ora.x_shape, generator=g)
        y = torch.randn(3, *ora.y_shape, generator=g)
        got = op.to(DEV).apply(x.to(DEV)).cpu()
        assert got.shape == (3, *ora.y_shape), name
        assert rel_err(got, ora.apply(x)) < 2e-6, (name, shape)
        got_t = op.apply_transpose(y.to(DEV)).cpu()
        assert got_t.shape == (3, *ora.x_shape), name
        assert rel_err(got_t, ora.adjoint(y)) < 2e-6, (name, shape)


def test_inpainting_mask_variants_and_roundtrip():
    """tests/operators/test_inpainting.py:15-57 of the reference, on the CUDA path."""
    from samplers_b200 import operators as pops
    shape = (3, 16, 16)
    for op in (pops.CenterInpaintingOperator(shape, 0.5), pops.CenterOutpaintingOperator(shape, 0.5),
               pops.SidePaintingOperator(shape, 0.25, left=False), pops.RandomInpaintingOperator(shape, 0.7)):
        op = op.to(DEV)
        m, n = op.shape
        assert n == 3 * 16 * 16 and op.get_singular_values().numel() == m and op.y_shape == (m,)
        x = torch.randn(2, *shape, device=DEV)
        back = op.apply_V(op.apply_V_transpose(x))
        keep = ~op.mask
        assert torch.equal(back[:, keep], x[:, keep]) and torch.all(back[:, op.mask] == 0)
        assert torch.equal(op.apply_pseudo_inverse(op.apply(x)), back)
    dense = pops.InpaintingOperator(shape, pops.get_mask_random(shape), flatten=False).to(DEV)
    x = torch.randn(2, *shape, device=DEV)
    assert torch.equal(dense.apply(x), x.masked_fill(dense.mask, 0))


def test_operator_autograd_backward_is_the_adjoint():
    from samplers_b200 import operators as pops
    shape = (3, 32, 32)
    op = pops.GaussianBlurOperator(shape, 9, 1.5).to(DEV)
    x = torch.randn(2, *shape, device=DEV, requires_grad=True)
    c = torch.randn(2, *shape, device=DEV)
    (op.apply(x) * c).sum().backward()
    assert rel_err(x.grad.cpu(), op.apply_transpose(c).cpu()) < 1e-6


FULL = (3, 256, 256)


@pytest.mark.parametrize("which", ["gblur61", "motion61", "box4", "mask"])
def test_full_size_adjoint_identity_and_linearity(which):
    """<A x, y> = <x, A^T y> and A(a x1 + x2) = a A x1 + A x2 at L=16, 3x256x256."""
    from samplers_b200 import operators as pops
    if which == "gblur61":
        op = pops.GaussianBlurOperator(FULL)
    elif which == "motion61":
        op = pops.MotionBlurOperator(FULL, kernel_size=61, angle_deg=37.0)
    elif which == "box4":
        op = pops.BoxDownsampleOperator(FULL, 4)
    else:
        op = pops.RandomInpaintingOperator(FULL, 0.7, flatten=False)
    op = op.to(DEV)
    g = torch.Generator(device=DEV).manual_seed(0)
    x1 = torch.randn(16, *op.x_shape, device=DEV, generator=g)
    x2 = torch.randn(16, *op.x_shape, device=DEV, generator=g)
    y = torch.randn(16, *op.y_shape, device=DEV, generator=g)
    ax1, aty = op.apply(x1), op.apply_transpose(y)
    lhs = (ax1.double() * y.double()).sum()
    rhs = (x1.double() * aty.double()).sum()
    assert abs(float(lhs - rhs)) < 1e-5 * max(1.0, abs(float(lhs)))
    lin = op.apply(0.5 * x1 + x2)
    assert rel_err(lin.cpu(), (0.5 * ax1 + op.apply(x2)).cpu()) < 2e-6


def test_full_size_blur_matches_oracle_arithmetic_on_device():
    """Oracle operator code executed on CUDA tensors as the checker (cuDNN fp32, TF32 off)."""
    from samplers_b200 import operators as pops
    torch.backends.cudnn.allow_tf32 = False
    op = pops.GaussianBlurOperator(FULL).to(DEV)
    ora = oops.OracleGaussianBlur(FULL, 61, 3.0)
    x = torch.randn(4, *FULL, device=DEV, generator=torch.Generator(device=DEV).manual_seed(1))
    assert rel_err(op.apply(x).cpu(), ora.apply(x).cpu()) < 2e-6
    assert rel_err(op.apply_transpose(x).cpu(), ora.adjoint(x).cpu()) < 2e-6


def test_full_size_k1_consistent_with_standalone_kernels():
    """K1 (fused) == w * A^T(y - A x0) / sa assembled from the stand-alone kernels, all five operator kinds."""
    from samplers_b200 import _native, operators as pops
    L = 16
    gen = torch.Generator(device=DEV).manual_seed(2)
    for op in (pops.IdentityOperator(FULL), pops.RandomInpaintingOperator(FULL, 0.7, flatten=False),
               pops.BoxDownsampleOperator(FULL, 4), pops.GaussianBlurOperator(FULL),
               pops.MotionBlurOperator(FULL, kernel_size=61, angle_deg=20.0)):
        op = op.to(DEV)
        nat = op._native_cached(torch.device(DEV))
        x = torch.randn(L, nat.n, device=DEV, generator=gen)
        eps = torch.randn(L, nat.n, device=DEV, generator=gen)
        y = torch.randn(1, nat.n_y, device=DEV, generator=gen)
        if isinstance(op, pops.InpaintingOperator):  # dense observations are zero at missing pixels
            y = y * (~op.mask).flatten().float()
        sa, s1, w = 0.8, 0.6, 400.0
        cot = torch.empty_like(x)
        part = torch.empty(L, nat.err_parts, device=DEV)
        wsb = nat.workspace_bytes(L)
        ws = torch.empty(wsb // 4, device=DEV) if wsb else None
        x0 = torch.empty_like(x)
        _native.tweedie(x, eps, sa, s1, x0)
        _native.dps_pre(nat, x, eps, y, L, sa, s1, w, cot, part, ws)
        r = y - nat.apply(x0)
        ref = nat.adjoint(r.contiguous()) * w / sa
        assert rel_err(cot.cpu(), ref.cpu()) < 2e-6, type(op).__name__
        assert rel_err(part.sum(1).cpu(), r.double().square().sum(1).float().cpu()) < 1e-5, type(op).__name__


def test_cluster_fused_blur_k1_matches_three_launch_path(psx_env):
    """The opt-in single-launch cluster/DSMEM CUDA-core kernel (PSX_FUSED=1) against the three-launch strip kernels."""
    from samplers_b200 import _native, operators as pops
    op = pops.GaussianBlurOperator(FULL).to(DEV)
    nat = op._native_cached(torch.device(DEV))
    L = 5
    gen = torch.Generator(device=DEV).manual_seed(7)
    x = torch.randn(L, nat.n, device=DEV, generator=gen)
    eps = torch.randn(L, nat.n, device=DEV, generator=gen)
    y = torch.randn(1, nat.n_y, device=DEV, generator=gen)
    ws = torch.empty(nat.workspace_bytes(L) // 4, device=DEV)
    outs = []
    for fused in (False, True):
        psx_env(PSX_FUSED="1" if fused else None, PSX_NO_TC="1")
        cot, part = torch.empty_like(x), torch.empty(L, nat.err_parts, device=DEV)
        _native.dps_pre(nat, x, eps, y, L, 0.8, 0.6, 400.0, cot, part, ws)
        outs.append((cot.clone(), part.sum(1).clone()))
    assert rel_err(outs[1][0].cpu(), outs[0][0].cpu()) < 2e-6
    assert rel_err(outs[1][1].cpu(), outs[0][1].cpu()) < 1e-6


@pytest.mark.parametrize("shape,taps", [((3, 512, 512), (61, 3.0)), ((1, 384, 512), (61, 3.0)), ((2, 512, 256), (61, 3.0)),
                                        ((3, 288, 320), (9, 1.5)), ((1, 512, 512), (21, 2.0))])
@pytest.mark.parametrize("L", [2, 5])
def test_blur_k1_on_planes_up_to_512_matches_oracle(shape, taps, L):
    """Configs 4-5 run the blur on 512 x 512 pixel planes: the strip kernels with 16-column strips and two TMA boxes
    per strip, against the oracle's torch expression executed on the device (cuDNN fp32, TF32 off)."""
    from samplers_b200 import _native, operators as pops
    torch.backends.cudnn.allow_tf32 = False
    op = pops.GaussianBlurOperator(shape, kernel_size=taps[0], sigma=taps[1]).to(DEV)
    ora = oops.OracleGaussianBlur(shape, taps[0], taps[1])
    nat = op._native_cached(torch.device(DEV))
    gen = torch.Generator(device=DEV).manual_seed(11)
    x = torch.randn(L, *shape, device=DEV, generator=gen)
    eps = torch.randn(L, *shape, device=DEV, generator=gen)
    y = torch.randn(1, *shape, device=DEV, generator=gen)
    sa, s1, w = 0.8, 0.6, 400.0
    cot = torch.empty(L, nat.n, device=DEV)
    part = torch.empty(L, nat.err_parts, device=DEV)
    ws = torch.empty(nat.workspace_bytes(L) // 4, device=DEV)
    _native.dps_pre(nat, x.view(L, -1), eps.view(L, -1), y.view(1, -1), L, sa, s1, w, cot, part, ws)
    x0 = (x - s1 * eps) / sa
    r = y - ora.apply(x0)
    ref = ora.adjoint(r) * (w / sa)
    assert rel_err(cot.cpu(), ref.reshape(L, -1).cpu()) < 5e-6
    assert rel_err(part.sum(1).cpu(), r.double().square().sum((1, 2, 3)).float().cpu()) < 1e-5
    assert rel_err(op.apply(x).cpu(), ora.apply(x).cpu()) < 2e-6
    assert rel_err(op.apply_transpose(x).cpu(), ora.adjoint(x).cpu()) < 2e-6


@pytest.mark.parametrize("env", [{}, {"PSX_SPLIT": "2"}, {"PSX_SPLIT": "4"}, {"PSX_NO_FAST16": "1"}, {"PSX_NO_PIPE": "1"},
                                 {"PSX_FUSED": "1"}, {"PSX_NO_TC": None}])
def test_every_blur_k1_code_path_gives_the_same_answer(psx_env, env):
    """The tensor-core kernel (the default for 256 x 256 planes; the last case), the strip kernels, the multi-group
    schedule, the 8-output pipelined kernels, the one-tile-per-CTA kernels and the cluster kernel are six
    implementations of one function.  The base is the strip-kernel path (PSX_NO_TC)."""
    from samplers_b200 import _native, operators as pops
    psx_env(PSX_SPLIT=None, PSX_NO_FAST16=None, PSX_NO_PIPE=None, PSX_FUSED=None, PSX_NO_TC="1")
    op = pops.GaussianBlurOperator(FULL).to(DEV)
    nat = op._native_cached(torch.device(DEV))
    L = 8
    gen = torch.Generator(device=DEV).manual_seed(3)
    x = torch.randn(L, nat.n, device=DEV, generator=gen)
    eps = torch.randn(L, nat.n, device=DEV, generator=gen)
    y = torch.randn(2, nat.n_y, device=DEV, generator=gen)          # two observations, 4 samples each
    ws = torch.empty(nat.workspace_bytes(L) // 4, device=DEV)

    def run():
        cot, part = torch.empty_like(x), torch.empty(L, nat.err_parts, device=DEV)
        _native.dps_pre(nat, x, eps, y, 4, 0.8, 0.6, 400.0, cot, part, ws)
        torch.cuda.synchronize()
        return cot, part.sum(1)

    base = run()
    psx_env(**env)
    got = run()
    if "PSX_SPLIT" in env:                  # same kernels, same tiles per sample: bit-identical
        assert torch.equal(got[0], base[0]) and torch.equal(got[1], base[1])
    else:
        assert rel_err(got[0].cpu(), base[0].cpu()) < 2e-6 and rel_err(got[1].cpu(), base[1].cpu()) < 1e-6


@pytest.mark.parametrize("form", ["rows", "cols"])
@pytest.mark.parametrize("psf", ["line20", "line80", "walk"])
@pytest.mark.parametrize("shape", [(2, 45, 70), (1, 64, 128)])
def test_motion_psf_row_and_column_segment_kernels_match_oracle(monkeypatch, form, psf, shape):
    """The 2-D PSF kernels in both segment forms (PSX_PSF_FORM forces one at descriptor creation): A, A^T against the
    oracle's conv2d, adjoint identity, and the fused K1 against the stand-alone kernels, ragged sizes included."""
    from samplers_b200 import _native, operators as pops
    from samplers_b200.operators.blur import motion_line_kernel, motion_walk_kernel
    torch.backends.cudnn.allow_tf32 = False
    monkeypatch.setenv("PSX_PSF_FORM", form)
    k2d = {"line20": lambda: motion_line_kernel(21, 20.0), "line80": lambda: motion_line_kernel(21, 80.0),
           "walk": lambda: motion_walk_kernel(21, 0.5, 3)}[psf]()
    op = pops.MotionBlurOperator(shape, kernel=k2d).to(DEV)
    ora = oops.OracleConv2dBlur(shape, k2d)
    nat = op._native_cached(torch.device(DEV))
    L = 3
    gen = torch.Generator(device=DEV).manual_seed(4)
    x = torch.randn(L, *shape, device=DEV, generator=gen)
    yv = torch.randn(L, *shape, device=DEV, generator=gen)
    ax, aty = op.apply(x), op.apply_transpose(yv)
    ora.kernel2d = ora.kernel2d.to(DEV)
    assert rel_err(ax.cpu(), ora.apply(x).cpu()) < 2e-6
    assert rel_err(aty.cpu(), ora.adjoint(yv).cpu()) < 2e-6
    lhs, rhs = (ax.double() * yv.double()).sum(), (x.double() * aty.double()).sum()
    assert abs(float(lhs - rhs)) < 1e-6 * max(1.0, abs(float(lhs)))
    eps = torch.randn(L, nat.n, device=DEV, generator=gen)
    y = torch.randn(1, nat.n_y, device=DEV, generator=gen)
    sa, s1, w = 0.8, 0.6, 400.0
    cot, part = torch.empty(L, nat.n, device=DEV), torch.empty(L, nat.err_parts, device=DEV)
    ws = torch.empty(nat.workspace_bytes(L) // 4, device=DEV)
    xf = x.reshape(L, -1).contiguous()
    _native.dps_pre(nat, xf, eps, y, L, sa, s1, w, cot, part, ws)
    x0 = torch.empty_like(xf)
    _native.tweedie(xf, eps, sa, s1, x0)
    r = y - nat.apply(x0)
    assert rel_err(cot.cpu(), (nat.adjoint(r.contiguous()) * w / sa).cpu()) < 2e-6
    assert rel_err(part.sum(1).cpu(), r.double().square().sum(1).float().cpu()) < 1e-5
