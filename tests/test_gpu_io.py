"""I/O row on the GPU, through the C ABI: bit-exact against the reference-recorded vectors and the
oracle on random data; observation simulation through InverseProblem.from_clean_data."""
import os

import numpy as np
import pytest
import torch

from oracle import io as oio
from oracle import operators as oops

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_image_kernels_match_reference_vectors():
    from samplers_b200.utils.image import pil_to_tensor, tensor_to_pil, tensor_to_uint8, uint8_to_tensor
    g = np.load(os.path.join(GOLD, "io_image.npz"))
    u8 = tensor_to_uint8(torch.from_numpy(g["x"]).to(DEV))
    assert u8.dtype == torch.uint8 and np.array_equal(u8.cpu().numpy(), g["u8"])
    back = uint8_to_tensor(torch.from_numpy(g["img"]).to(DEV))
    assert np.array_equal(back.cpu().numpy(), g["back"])
    assert np.array_equal(np.array(tensor_to_pil(torch.from_numpy(g["x"]).to(DEV))), g["u8"])
    from PIL import Image
    t = pil_to_tensor(Image.fromarray(g["img"]), device=DEV)
    assert t.is_cuda and np.array_equal(t.cpu().numpy(), g["back"])


@pytest.mark.parametrize("shape", [(16, 3, 256, 256), (2, 5, 3, 33, 17), (1, 1, 7, 9), (4, 64, 64)])
def test_image_kernels_match_oracle_on_random_batches(shape):
    from samplers_b200.utils.image import tensor_to_uint8, uint8_to_tensor
    gen = torch.Generator().manual_seed(5)
    x = torch.randn(shape, generator=gen) * 0.7
    u8 = tensor_to_uint8(x.to(DEV))
    assert torch.equal(u8.cpu(), oio.image_to_u8(x))
    f = uint8_to_tensor(u8)
    assert torch.equal(f.cpu(), oio.image_from_u8(u8.cpu()))
    d = u8.cpu().int() - tensor_to_uint8(f).cpu().int()   # .byte() truncates: a byte may come back one lower
    assert int(d.min()) >= 0 and int(d.max()) <= 1


@pytest.mark.parametrize("kind", ["gaussian", "poisson"])
def test_observe_matches_reference_vectors(kind):
    from samplers_b200 import _native
    from samplers_b200.operators import IdentityOperator
    g = np.load(os.path.join(GOLD, "io_observation.npz"))
    x = torch.from_numpy(g["x_true"]).to(DEV)
    raw = torch.from_numpy(g[f"{kind}_raw"]).to(DEV)
    param = float(g[f"{kind}_param"])
    a, b = (param, 0.0) if kind == "gaussian" else (1.0, -param)
    nat = IdentityOperator((3, 8, 8))._native_cached(torch.device(DEV))
    y = torch.empty_like(x)
    _native.observe(nat, x.reshape(2, -1), raw.reshape(2, -1), a, b, y.view(2, -1))
    assert np.array_equal(y.cpu().numpy(), g[f"{kind}_y"])


@pytest.mark.parametrize("op_name", ["identity", "blur", "box", "mask_flat"])
@pytest.mark.parametrize("kind", ["gaussian", "poisson"])
def test_from_clean_data_equals_oracle_composition(op_name, kind):
    from samplers_b200 import operators as P
    from samplers_b200.inverse_problem import InverseProblem
    from samplers_b200.noise import GaussianNoise, PoissonNoise
    shape = (3, 32, 32)
    mask = torch.rand(shape, generator=torch.Generator().manual_seed(1)) < 0.7
    op, ora = {
        "identity": (lambda: P.IdentityOperator(shape), lambda: oops.OracleIdentity(shape)),
        "blur": (lambda: P.GaussianBlurOperator(shape, 9, 1.5), lambda: oops.OracleGaussianBlur(shape, 9, 1.5)),
        "box": (lambda: P.BoxDownsampleOperator(shape, 4), lambda: oops.OracleBoxDownsample(shape, 4)),
        "mask_flat": (lambda: P.InpaintingOperator(shape, mask, flatten=True), lambda: oops.OracleMaskGather(shape, mask)),
    }[op_name]
    op, ora = op().to(DEV), ora()
    noise, param = (GaussianNoise(sigma=0.05), 0.05) if kind == "gaussian" else (PoissonNoise(rate=4.0), 4.0)
    x = torch.rand((3, *shape), generator=torch.Generator().manual_seed(2)) * 2 - 1
    prob = InverseProblem.from_clean_data(x.to(DEV), operator=op, noise=noise,
                                          rng=torch.Generator(device=DEV).manual_seed(9))
    rng = torch.Generator(device=DEV).manual_seed(9)
    yshape = prob.observation.shape
    raw = (torch.randn(yshape, device=DEV, generator=rng) if kind == "gaussian"
           else torch.poisson(torch.full(yshape, param, device=DEV), generator=rng))
    # the operator pass itself is compared to 2e-6 elsewhere; here: same clean signal, noise folded in bit-exactly
    clean = op(x.to(DEV))
    want = oio.simulate_observation(clean.cpu(), kind, param, raw.cpu())
    assert prob.observation.shape == want.shape and torch.equal(prob.observation.cpu(), want)
    ref_clean = ora.apply(x).reshape(want.shape)
    assert float((clean.cpu() - ref_clean).abs().max()) < 2e-6 * max(1.0, float(ref_clean.abs().max()))
