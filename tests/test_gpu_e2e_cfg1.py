"""BASELINE config 1 end to end (the reference's own demo call, scripts/run_dps.py:29-39): DPSSampler, identity
operator, sigma = 0.05, 3 x 64 x 64, 50 sampling steps, one reconstruction, the ddpm-celebahq-256 UNet at random init
(seed 1234) -- the whole `sampler(problem)` on the GPU against the recording of the unmodified reference on CPU
(oracle/make_golden_cfg1.py), same observation, same noise draws.  north_star: restored images within 0.05 dB PSNR.

Which gamma.  At gamma = sigma^2 the run is well conditioned and the GPU result equals the recording (measured: PSNR
difference 0.00000 dB, relative error of the final estimate 1.9e-5).  At the default gamma = 1 every guided step
displaces the state by ~400 / sqrt(acp_t) in norm (hard part 2) and the iteration is chaotic: the UNMODIFIED
reference does not reproduce ITSELF -- the same CPU run with 3 instead of 4 host threads (a different fp32 reduction
order inside ATen) parts from the recording by 1e-3 at guided step 12, 5e-2 at 24, 0.45 at 36 and ends at 26.05 dB
instead of 5.80 dB.  No implementation can meet 0.05 dB there; for gamma = 1 this test therefore checks what IS
defined: identical start, agreement while the perturbation is still small (step 12), a finite result -- and prints the
final difference."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import dps as odps
from tests._golden import rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "cfg1_e2e.npz")


def _run(gamma, shape, steps, sigma, y, states=None):
    from samplers_b200.inverse_problem import InverseProblem
    from samplers_b200.networks import DDPMNetwork
    from samplers_b200.noise import GaussianNoise
    from samplers_b200.operators import IdentityOperator
    from samplers_b200.samplers import DPSSampler
    net = DDPMNetwork.from_config("google/ddpm-celebahq-256", seed=1234, device=DEV)
    if states is not None:                         # the state every network call sees
        inner = net.forward

        def recording_forward(sample, t):
            states.append(sample.detach().clone().cpu())
            return inner(sample, t)
        net.forward = recording_forward
    prob = InverseProblem(operator=IdentityOperator(x_shape=shape).to(DEV), observation=y.to(DEV),
                          noise=GaussianNoise(sigma=sigma))
    gz = torch.Generator().manual_seed(2)          # the recording's generator, drawn in the sampler's own order
    sampler = DPSSampler(net)
    sampler.draw = lambda shp, device, dtype: torch.randn(shp, generator=gz).to(device)
    return sampler(prob, num_sampling_steps=steps, num_reconstructions=1, gamma=gamma, eta=1.0).cpu()


@pytest.mark.parametrize("tag", ["gamma_sigma2", "gamma1"])
def test_config1_whole_sampler_matches_reference_psnr(tag):
    g = np.load(PATH)
    meta = json.loads(bytes(g["meta"]).decode())
    tf32 = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False   # fp32 network, as on the CPU
    states = []
    try:
        out = _run(meta["gammas"][tag], tuple(meta["shape"]), meta["steps"], meta["sigma"], torch.from_numpy(g["y"]),
                   states)
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = tf32
    ref, x_true = torch.from_numpy(g[f"x0_{tag}"]), torch.from_numpy(g["x_true"])
    assert out.shape == ref.shape and torch.isfinite(out).all()
    d_psnr = abs(odps.psnr(out, x_true) - meta[f"psnr_{tag}"])
    print(f"cfg1 {tag}: PSNR ours {odps.psnr(out, x_true):.4f} dB, reference {meta[f'psnr_{tag}']:.4f} dB, "
          f"|diff| {d_psnr:.5f} dB, relative error of the final estimate {rel_err(out, ref):.3e}")
    kept = torch.from_numpy(g[f"states_{tag}"])
    errs = [rel_err(states[k].reshape(kept[i].shape), kept[i]) for i, k in enumerate(meta["keep"])]
    print(f"cfg1 {tag}: relative error of the state entering guided steps {meta['keep']}: "
          + ", ".join(f"{e:.2e}" for e in errs))
    assert errs[0] < 1e-6                                  # the same start
    if tag == "gamma_sigma2":
        assert max(errs) < 1e-4 and d_psnr < 0.05          # north_star's end-to-end bar
    else:
        assert errs[1] < 5e-2                              # still together where the reference agrees with itself
