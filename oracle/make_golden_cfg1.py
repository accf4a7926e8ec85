"""Record BASELINE config 1 END TO END from the UNMODIFIED reference (build container only; TEST INFRA).

    python -m oracle.make_golden_cfg1        ->  tests/golden/cfg1_e2e.npz

The reference's own demo call (scripts/run_dps.py:29-39): DPSSampler, IdentityOperator, GaussianNoise(sigma = 0.05),
x 3 x 64 x 64, 50 sampling steps (48 guided), one reconstruction, eta = 1 -- with the ddpm-celebahq-256 UNet
architecture at random init (torch.manual_seed(1234), the weights DDPMNetwork.from_config builds) on CPU fp32, the
noise drawn from torch.Generator(seed 2) in the sampler's own order.  Two guidance scales: the default gamma = 1 (the
chaotic regime of SURVEY 7 hard part 2: a 1e-7 difference grows by orders of magnitude over 48 steps) and the
well-conditioned gamma = sigma^2.  Only seeds, the observation and the final estimates are stored; the GPU test
(tests/test_gpu_e2e_cfg1.py) redraws the same noise from the same generator.

Not part of tests/test_golden_reproducible.py: 96 forward + backward passes of a 114 M-parameter UNet take minutes on
the host cores, and with more than one thread the fp32 reduction order (hence the last bits) is not fixed.
"""
from __future__ import annotations

import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import dps as odps, ref_shim  # noqa: E402
from oracle.schedule import ddpm_linear_alphas_cumprod  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
SHAPE, STEPS, SIGMA = (3, 64, 64), 50, 0.05
GAMMAS = {"gamma1": 1.0, "gamma_sigma2": SIGMA * SIGMA}
KEEP = (0, 12, 24, 36, 47)     # guided-step indices whose incoming state is kept (where do the runs part?)


class _Eps(torch.nn.Module):
    def __init__(self, unet):
        super().__init__()
        self.unet = unet

    def forward(self, x, t):
        return self.unet(sample=x, timestep=t).sample


def main():
    from samplers_b200.networks.unet2d import CELEBAHQ_256, UNet2DModel   # the network stays in torch on both sides
    ref = ref_shim.load_reference()
    torch.manual_seed(1234)
    unet = UNet2DModel(**CELEBAHQ_256).eval().requires_grad_(False)
    x_true = torch.rand(SHAPE, generator=torch.Generator().manual_seed(0)) * 2 - 1
    arrays, meta = {}, dict(shape=SHAPE, steps=STEPS, sigma=SIGMA, gammas=GAMMAS, keep=list(KEEP), seed_net=1234,
                            seed_x=0, seed_y=1, seed_z=2, torch=torch.__version__, threads=torch.get_num_threads())
    for tag, gamma in GAMMAS.items():
        rec = ref_shim.RecordingNet(_Eps(unet))
        network = ref.networks.DDPMNetwork(ref_shim.FakeDDPMPipeline(rec, ddpm_linear_alphas_cumprod(), 1000))
        problem = ref.inverse_problem.InverseProblem.from_clean_data(
            x_true, operator=ref.operators.IdentityOperator(x_shape=SHAPE), noise=ref.noise.GaussianNoise(sigma=SIGMA),
            rng=torch.Generator().manual_seed(1))
        gz = torch.Generator().manual_seed(2)
        t0 = time.perf_counter()
        with ref_shim.injected_noise(lambda shape_: ref_shim.REAL_RANDN(shape_, generator=gz)):
            out = ref.samplers.DPSSampler(network)(problem, num_sampling_steps=STEPS, num_reconstructions=1,
                                                  gamma=gamma, eta=1.0)
        wall = time.perf_counter() - t0
        assert out.shape == SHAPE and len(rec.calls) == STEPS - 1, (out.shape, len(rec.calls))
        arrays[f"x0_{tag}"] = out.numpy()
        arrays[f"states_{tag}"] = np.stack([rec.calls[k]["x_t"].numpy() for k in KEEP])
        meta[f"psnr_{tag}"] = float(odps.psnr(out, x_true))
        meta[f"wall_s_{tag}"] = wall
        arrays["y"] = problem.observation.numpy()
        print(tag, "gamma", gamma, "psnr", meta[f"psnr_{tag}"], "wall", round(wall, 1), "s", flush=True)
    arrays["x_true"] = x_true.numpy()
    arrays["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    os.makedirs(OUT, exist_ok=True)
    np.savez_compressed(os.path.join(OUT, "cfg1_e2e.npz"), **arrays)


if __name__ == "__main__":
    main()
