"""Oracle PSLD restatement vs recordings of the UNMODIFIED reference PSLDSampler (CPU)."""
import pytest
import torch

from oracle import psld as opsld
from tests._golden import PsldGolden, psld_names, rel_err

TOL = 1e-5


@pytest.mark.parametrize("name", psld_names())
def test_psld_teacher_forced(name):
    g = PsldGolden(name)
    core, op, m = g.core(), g.oracle_op(), g.meta
    y_flat = g.y_flat(op)
    aty = op.adjoint(y_flat)
    worst = 0.0
    for k in range(g.K):
        out = opsld.psld_step(core.eps, core.decode, core.encode, g["z_t"][k], t=m["t"][k], t_prev=m["t_prev"][k],
                              s=m["s"], acp=g["acp"], op=op, y_flat=y_flat, aty_flat=aty, omega=m["omega"],
                              gamma=m["gamma"], eta=m["eta"], noise=g["noise"][k])
        assert rel_err(out["grad"], g["grad"][k]) < TOL
        worst = max(worst, rel_err(out["z_next"], g["z_next"][k]))
    assert worst < TOL, worst


@pytest.mark.parametrize("name", psld_names())
def test_psld_free_running(name):
    g = PsldGolden(name)
    core, op, m = g.core(), g.oracle_op(), g.meta
    draws = iter([g["z_init"]] + [g["noise"][k] for k in range(g.K)])
    out = opsld.psld_sample(core.eps, core.decode, core.encode, acp=g["acp"], timesteps=g["timesteps"].tolist(),
                            op=op, y_flat=g.y_flat(op), latent_shape=tuple(m["latent_shape"]), leading=g.L,
                            omega=m["omega"], gamma=m["gamma"], eta=m["eta"], draw=lambda s: next(draws))
    assert rel_err(out, g["x_out"].reshape(out.shape)) < 1e-4
