"""Oracle ReSample restatement vs recordings of the UNMODIFIED reference ReSampleSampler (CPU)."""
import pytest
import torch

from oracle import resample as ors
from tests._golden import ResampleGolden, rel_err, resample_names


@pytest.mark.parametrize("name", resample_names())
def test_resample_free_running_matches_reference(name):
    g = ResampleGolden(name)
    core, op, m = g.core(), g.oracle_op(), g.meta
    draws = iter(g["draws"])
    calls = []

    def eps_fn(z, t):
        calls.append((z.detach().clone(), int(t)))
        return core.eps(z, t)

    out = ors.resample_sample(eps_fn, core.decode, core.encode, acp=g["acp"], timesteps=g["timesteps"].tolist(),
                              op=op, y_flat=g.y_flat(op), latent_shape=tuple(m["latent_shape"]), leading=g.L,
                              eps=g.eps_threshold(), draw=lambda s: next(draws), **m["kw"])
    assert len(calls) == m["n_net_calls"]
    assert [t for _, t in calls] == g["call_t"].tolist()          # same control flow (time travel, stages)
    for (z, _), ref in zip(calls, g["call_z"]):
        assert rel_err(z, ref) < 1e-4
    assert rel_err(out, g["x_out"].reshape(out.shape)) < 1e-4
    assert next(draws, None) is None                               # consumed exactly the reference's draws


def test_stochastic_resample_and_sigma_formulas():
    """tests/samplers/test_resample.py:188-197 of the reference + values."""
    b = 2
    p, x, n = torch.randn(b, 3, 4, 4), torch.randn(b, 3, 4, 4), torch.randn(b, 3, 4, 4)
    a, s = torch.full((b, 1, 1, 1), 0.8), torch.full((b, 1, 1, 1), 0.5)
    out = ors.stochastic_resample(p, x, a, s, n)
    assert out.shape == p.shape
    ref = (0.5 * 0.8 ** 0.5 * p + 0.2 * x) / 0.7 + n * (1 / (2 + 5)) ** 0.5
    assert torch.allclose(out, ref, atol=1e-6)
    assert torch.allclose(ors.compute_sigma(40.0, torch.tensor(0.5), torch.tensor(0.8)),
                          torch.tensor(40.0 * 0.2 / 0.5 * (1 - 0.5 / 0.8)))
