"""Records golden vectors for the I/O row from the UNMODIFIED reference:
samplers.utils.image.tensor_to_pil / pil_to_tensor and InverseProblem.from_clean_data.

    python -m oracle.make_golden_io      (needs /root/reference; writes tests/golden/io_*.npz)
"""
from __future__ import annotations

import os

import numpy as np
import torch

from . import io as oio
from .ref_shim import load_reference

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def main():
    load_reference()
    from samplers.inverse_problem import InverseProblem
    from samplers.noise import GaussianNoise, PoissonNoise
    from samplers.operators.identity import IdentityOperator
    from samplers.utils.image import pil_to_tensor, tensor_to_pil
    from PIL import Image

    # ---- images: random values beyond [-1, 1], the exact byte grid, and values that sit on rounding edges
    g = torch.Generator().manual_seed(11)
    x = torch.rand((3, 24, 20), generator=g) * 2.6 - 1.3
    grid = (torch.arange(256, dtype=torch.float32) / 255) * 2 - 1
    x[0, :12, :20].reshape(-1)[:240] = grid[:240]
    x[1, 0, :16] = grid[240:]
    x[2, 0, :8] = torch.tensor([-1.0, 1.0, 0.0, -0.0, 0.99999994, -0.99999994, 0.5, -0.5])
    u8 = np.array(tensor_to_pil(x))                                   # (H, W, C) uint8
    all_bytes = np.arange(256, dtype=np.uint8).reshape(16, 16)
    img = np.stack([all_bytes, all_bytes[::-1], all_bytes.T], axis=-1)  # every byte value in every channel
    back = pil_to_tensor(Image.fromarray(img)).numpy()                # (C, H, W) float32
    assert np.array_equal(oio.image_to_u8(x).numpy(), u8), "oracle image_to_u8 differs from the reference"
    assert np.array_equal(oio.image_from_u8(torch.from_numpy(img)).numpy(), back), "oracle image_from_u8 differs"
    np.savez_compressed(os.path.join(OUT, "io_image.npz"), x=x.numpy(), u8=u8, img=img, back=back)

    # ---- observation simulation through the reference's from_clean_data (identity operator, both noise models)
    x_true = torch.rand((2, 3, 8, 8), generator=torch.Generator().manual_seed(12)) * 2 - 1
    op = IdentityOperator(x_shape=(3, 8, 8))
    out = {"x_true": x_true.numpy()}
    for kind, noise, param in (("gaussian", GaussianNoise(sigma=0.05), 0.05), ("poisson", PoissonNoise(rate=4.0), 4.0)):
        prob = InverseProblem.from_clean_data(x_true, operator=op, noise=noise, rng=torch.Generator().manual_seed(13))
        rng = torch.Generator().manual_seed(13)
        raw = (torch.randn(x_true.shape, generator=rng) if kind == "gaussian"
               else torch.poisson(torch.full(x_true.shape, param), generator=rng))
        y = prob.observation
        assert torch.equal(oio.simulate_observation(x_true, kind, param, raw), y), f"oracle {kind} observation differs"
        out[f"{kind}_raw"], out[f"{kind}_y"], out[f"{kind}_param"] = raw.numpy(), y.numpy(), np.float32(param)
    np.savez_compressed(os.path.join(OUT, "io_observation.npz"), **out)
    print("wrote io_image.npz, io_observation.npz")


if __name__ == "__main__":
    main()
