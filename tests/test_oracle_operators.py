"""Adjoint (dot-product) tests for the oracle operators -- the specification the
CUDA operator kernels are held to.  CPU only."""
import pytest
import torch

from oracle import operators as oops


def _ops():
    shape = (3, 20, 28)
    g = torch.Generator().manual_seed(3)
    mask = torch.rand(shape, generator=g) < 0.7
    return {
        "identity": oops.OracleIdentity(shape),
        "identity_flat": oops.OracleIdentity(shape, flatten=True),
        "mask": oops.OracleMaskGather(shape, mask),
        "gblur9": oops.OracleGaussianBlur(shape, 9, 1.5),
        "gblur61": oops.OracleGaussianBlur(shape, 61, 3.0),
        "sep_asym": oops.OracleSeparableBlur(shape, torch.rand(7, generator=g), torch.rand(5, generator=g)),
        "motion_line": oops.OracleConv2dBlur(shape, oops.motion_line_kernel(11, 30.0)),
        "motion_walk": oops.OracleConv2dBlur(shape, oops.motion_walk_kernel(13, 0.5, seed=1)),
        "box4": oops.OracleBoxDownsample(shape, 4),
        "box2": oops.OracleBoxDownsample(shape, 2),
        "maskbox4": oops.OracleMaskedBox(shape, 4, (torch.rand(3, 5, 7, generator=g) >= 0.7).float()),
    }


@pytest.mark.parametrize("name", list(_ops().keys()))
def test_adjoint_identity(name):
    op = _ops()[name]
    g = torch.Generator().manual_seed(11)
    x = torch.randn(2, *op.x_shape, generator=g, dtype=torch.float64)
    y = torch.randn(2, *op.y_shape, generator=g, dtype=torch.float64)
    if hasattr(op, "taps_h"):
        op.taps_h, op.taps_v = op.taps_h.double(), op.taps_v.double()
    if hasattr(op, "kernel2d"):
        op.kernel2d = op.kernel2d.double()
    lhs = (op.apply(x) * y).sum()
    rhs = (x * op.adjoint(y)).sum()
    assert abs(float(lhs - rhs)) < 1e-9 * max(1.0, abs(float(lhs)))


def test_gaussian_taps_normalised_and_symmetric():
    w = oops.gaussian_taps(61, 3.0)
    assert w.numel() == 61 and abs(float(w.double().sum()) - 1) < 1e-6
    assert torch.equal(w, w.flip(0))


def test_separable_equals_outer_product_2d():
    shape = (3, 24, 24)
    sep = oops.OracleGaussianBlur(shape, 9, 1.5)
    full = oops.OracleConv2dBlur(shape, torch.outer(sep.taps_v, sep.taps_h))
    x = torch.randn(2, *shape, generator=torch.Generator().manual_seed(0))
    assert torch.allclose(sep.apply(x), full.apply(x), atol=1e-6)


def test_motion_kernels_sum_to_one():
    assert abs(float(oops.motion_line_kernel(61, 45.0).sum()) - 1) < 1e-5
    assert abs(float(oops.motion_walk_kernel(61, 0.5).sum()) - 1) < 1e-5


def test_masked_box_is_mask_after_box_and_has_its_pseudo_inverse():
    shape = (3, 16, 24)
    g = torch.Generator().manual_seed(5)
    keep = (torch.rand(3, 4, 6, generator=g) >= 0.7).double()
    op, box = oops.OracleMaskedBox(shape, 4, keep), oops.OracleBoxDownsample(shape, 4)
    x = torch.randn(2, *shape, generator=g, dtype=torch.float64)
    y = op.apply(x)
    assert torch.equal(y, box.apply(x) * keep) and torch.all(y[:, keep == 0] == 0)
    assert torch.allclose(op.apply(op.pinv(y)), y, atol=1e-12)          # A A^+ y = y on the range of A


def test_oracle_blur_box_and_motion_against_independent_implementations():
    """The blur / motion / box operators have no counterpart in the reference, so the oracle is their definition;
    here it is cross-checked against implementations that share no code with it: scipy.ndimage (separable and 2-D
    zero-padded correlation), a numpy reshape-mean (box) and the closed form of the sampled Gaussian."""
    import numpy as np
    from scipy import ndimage
    shape = (2, 20, 28)
    g = torch.Generator().manual_seed(9)
    x = torch.randn(1, *shape, generator=g, dtype=torch.float64)
    # taps: exp(-i^2 / (2 sigma^2)) normalised to sum 1
    w = oops.gaussian_taps(9, 1.5).double().numpy()
    i = np.arange(9) - 4
    ref_w = np.exp(-i ** 2 / (2 * 1.5 ** 2))
    assert np.allclose(w, ref_w / ref_w.sum(), atol=1e-7)
    # separable blur: rows with taps_h, then columns with taps_v (asymmetric taps on purpose: correlation, not convolution)
    th, tv = torch.rand(7, generator=g, dtype=torch.float64), torch.rand(5, generator=g, dtype=torch.float64)
    sep = oops.OracleSeparableBlur(shape, th, tv)
    sep.taps_h, sep.taps_v = th, tv
    got = sep.apply(x)[0].numpy()
    for c in range(shape[0]):
        r = ndimage.correlate1d(x[0, c].numpy(), th.numpy(), axis=1, mode="constant", cval=0.0)
        r = ndimage.correlate1d(r, tv.numpy(), axis=0, mode="constant", cval=0.0)
        assert np.allclose(got[c], r, atol=1e-12)
    # adjoint = correlation with the flipped taps = convolution
    gt = sep.adjoint(x)[0].numpy()
    for c in range(shape[0]):
        r = ndimage.convolve1d(x[0, c].numpy(), tv.numpy(), axis=0, mode="constant", cval=0.0)
        r = ndimage.convolve1d(r, th.numpy(), axis=1, mode="constant", cval=0.0)
        assert np.allclose(gt[c], r, atol=1e-12)
    # 2-D PSF
    k2 = oops.motion_line_kernel(11, 30.0).double()
    op2 = oops.OracleConv2dBlur(shape, k2)
    op2.kernel2d = k2
    got2 = op2.apply(x)[0].numpy()
    for c in range(shape[0]):
        assert np.allclose(got2[c], ndimage.correlate(x[0, c].numpy(), k2.numpy(), mode="constant", cval=0.0), atol=1e-12)
    # box: mean over non-overlapping 4 x 4 blocks
    box = oops.OracleBoxDownsample(shape, 4).apply(x)[0].numpy()
    ref = x[0].numpy().reshape(shape[0], shape[1] // 4, 4, shape[2] // 4, 4).mean(axis=(2, 4))
    assert np.allclose(box, ref, atol=1e-12)
