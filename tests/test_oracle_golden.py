"""The oracle restatement vs recordings of the UNMODIFIED reference
(tests/golden/*.npz, made by oracle/make_golden.py).  CPU only."""
import pytest
import torch

from oracle import dps as odps
from tests._golden import Golden, golden_names, rel_err

TOL = 1e-5  # north_star: per-step state within 1e-5 relative in fp32


@pytest.mark.parametrize("name", golden_names())
@pytest.mark.parametrize("form", ["autograd", "closed"])
def test_teacher_forced_steps(name, form):
    g = Golden(name)
    net, op, m = g.net(), g.oracle_op(), g.meta
    step = odps.dps_step_autograd if form == "autograd" else odps.dps_step_closed_form
    nk, nparam = m["noise"]
    worst = 0.0
    for k in range(g.K):
        out = step(net, g["x_t"][k], t=m["t"][k], t_prev=m["t_prev"][k], s=m["s"], acp=g["acp"],
                   op=op, y=g.y_flat(), noise_kind=nk, noise_param=torch.tensor(nparam),
                   gamma=m["gamma"], eta=m["eta"], z=g["z"][k])
        assert rel_err(out["eps"], g["eps"][k]) < 1e-6
        worst = max(worst, rel_err(out["x_next"], g["x_next"][k]))
        assert rel_err(out["grad"], g["grad"][k]) < TOL
    assert worst < TOL, worst


@pytest.mark.parametrize("name", golden_names())
def test_closed_form_pieces_match_recorded(name):
    """d, v, e2 stored in the fixture reproduce the reference's autograd gradient."""
    g = Golden(name)
    m = g.meta
    for k in range(g.K):
        s1 = (1 - g["acp"][m["t"][k]]) ** 0.5
        grad = g["d"][k] - s1 * g["v"][k]
        assert rel_err(grad, g["grad"][k]) < TOL


@pytest.mark.parametrize("name", ["identity_mock_eta05", "identity_ddim", "box4_gauss"])
def test_free_running_matches_reference_output(name):
    """Whole-sampler replay with the same injected noise (well-conditioned cases)."""
    g = Golden(name)
    m = g.meta
    draws = [g["x_init"]] + [g["z"][k] for k in range(g.K)]
    it = iter(draws)
    nk, nparam = m["noise"]
    out = odps.dps_sample(g.net(), acp=g["acp"], timesteps=g["timesteps"].tolist(), op=g.oracle_op(),
                          y=g.y_flat(), noise_kind=nk, noise_param=torch.tensor(nparam), leading=g.L,
                          gamma=m["gamma"], eta=m["eta"], draw=lambda shp: next(it))
    ref = g["x0_final"].reshape(out.shape)
    assert rel_err(out, ref) < 1e-3
    assert abs(odps.psnr(out, g["x_true"].expand_as(out).reshape(out.shape) if g["x_true"].numel() != out.numel()
                         else g["x_true"].reshape(out.shape)) -
               odps.psnr(ref, g["x_true"].expand_as(ref) if g["x_true"].numel() != ref.numel()
                         else g["x_true"].reshape(ref.shape))) < 0.05


def test_mock_schedule_has_nonzero_s():
    g = Golden("identity_mock_eta05")
    assert g.meta["s"] == 1 and float(g["acp"][1]) < 1.0
