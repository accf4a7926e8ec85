"""conv2d_rowseg2 (2-D row-segment PSFs on packed FFMA2, taps in the parameter bank) against the scalar-FFMA kernel it
replaces (PSX_NO_C2V2=1): the two evaluate every output with the same taps in the same order -- one FFMA2 is two FMAs --
so the residual and K1's cotangent must agree bit for bit, on full tiles and on ragged borders; the |r|^2 partial sums
add the same squares in a different thread order (8 x 2 outputs per thread, another row pairing) and agree to rounding."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _k1(nat, x, e, y, L):
    from samplers_b200 import _native
    cot = torch.full_like(x, float("nan"))
    part = torch.empty(L, nat.err_parts, device=DEV)
    ws = torch.empty(max(nat.workspace_bytes(L) // 4, 1), device=DEV)
    _native.dps_pre(nat, x, e, y, L, 0.8, 0.6, 400.0, cot, part, ws)
    torch.cuda.synchronize()
    return cot, part


@pytest.mark.parametrize("shape,angle", [((3, 256, 256), 30.0), ((1, 72, 100), 30.0), ((2, 64, 64), 10.0),
                                          ((1, 200, 36), 40.0)])
def test_ffma2_kernel_is_bitwise_the_scalar_kernel(shape, angle):
    from samplers_b200 import _native, operators as pops
    os.environ["PSX_PSF_FORM"] = "rows"   # the form conv2d_rowseg2 takes (read when the operator is created)
    try:
        op = pops.MotionBlurOperator(shape, kernel_size=61, angle_deg=angle).to(DEV)
        nat = op._native_cached(torch.device(DEV))
    finally:
        del os.environ["PSX_PSF_FORM"]
    L, n = 3, nat.n
    g = torch.Generator(device=DEV).manual_seed(1)
    x, e = (torch.randn(L, n, device=DEV, generator=g) for _ in range(2))
    y = torch.randn(1, nat.n_y, device=DEV, generator=g)
    new = _k1(nat, x, e, y, L)
    os.environ["PSX_NO_C2V2"] = "1"
    _native.load().psx_reload_env()
    try:
        old = _k1(nat, x, e, y, L)
    finally:
        del os.environ["PSX_NO_C2V2"]
        _native.load().psx_reload_env()
    assert torch.isfinite(new[0]).all()
    assert torch.equal(new[0], old[0])
    tot_new, tot_old = new[1].double().sum(1), old[1].double().sum(1)
    assert float(((tot_new - tot_old).abs() / tot_old).max()) < 1e-6
