"""I/O row, CPU: the oracle against vectors recorded from the unmodified reference
(oracle/make_golden_io.py), plus the host-side mirrors of samplers/utils/image.py."""
import os

import numpy as np
import pytest
import torch

from oracle import io as oio

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_image_oracle_matches_reference_vectors():
    g = np.load(os.path.join(GOLD, "io_image.npz"))
    assert np.array_equal(oio.image_to_u8(torch.from_numpy(g["x"])).numpy(), g["u8"])
    assert np.array_equal(oio.image_from_u8(torch.from_numpy(g["img"])).numpy(), g["back"])
    # u8 -> [-1,1] -> u8: torchvision's .byte() truncates, so a byte may come back one lower, never further off
    rt = oio.image_to_u8(oio.image_from_u8(torch.from_numpy(g["img"]))).numpy().astype(np.int32)
    d = g["img"].astype(np.int32) - rt
    assert d.min() >= 0 and d.max() <= 1


@pytest.mark.parametrize("kind", ["gaussian", "poisson"])
def test_observation_oracle_matches_reference_vectors(kind):
    g = np.load(os.path.join(GOLD, "io_observation.npz"))
    y = oio.simulate_observation(torch.from_numpy(g["x_true"]), kind, float(g[f"{kind}_param"]),
                                 torch.from_numpy(g[f"{kind}_raw"]))
    assert np.array_equal(y.numpy(), g[f"{kind}_y"])


def test_host_image_helpers_follow_the_reference_arithmetic():
    from samplers_b200.utils.image import pil_to_tensor, tensor_to_pil
    g = np.load(os.path.join(GOLD, "io_image.npz"))
    pil = tensor_to_pil(torch.from_numpy(g["x"]))
    assert pil.mode == "RGB" and np.array_equal(np.array(pil), g["u8"])
    from PIL import Image
    back = pil_to_tensor(Image.fromarray(g["img"]))
    assert back.dtype == torch.float32 and np.array_equal(back.numpy(), g["back"])
    with pytest.raises(ValueError):
        tensor_to_pil(torch.zeros(2, 3, 4, 4))       # one image at a time, as the reference


def test_noise_draw_affine_equals_sample():
    from samplers_b200.noise import GaussianNoise, PoissonNoise
    for noise in (GaussianNoise(sigma=0.05), PoissonNoise(rate=4.0)):
        raw, a, b = noise._draw_affine((2, 3, 5), "cpu", torch.Generator().manual_seed(3))
        want = noise.sample((2, 3, 5), generator=torch.Generator().manual_seed(3))
        assert torch.equal(torch.tensor(a) * raw + torch.tensor(b), want)
