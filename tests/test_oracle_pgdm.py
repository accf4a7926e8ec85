"""Oracle PGDM restatement vs recordings of the UNMODIFIED reference PGDMSampler (CPU)."""
import pytest

from oracle import pgdm as opg
from tests._golden import PgdmGolden, pgdm_names, rel_err


@pytest.mark.parametrize("name", pgdm_names())
def test_pgdm_teacher_forced(name):
    g = PgdmGolden(name)
    net, op, m = g.net(), g.oracle_op(), g.meta
    for k in range(g.K):
        out = opg.pgdm_step(net, g["x_t"][k], t=m["t"][k], t_prev=m["t_prev"][k], s=m["s"], acp=g["acp"], op=op,
                            y_flat=g.y_flat(), guidance_weight=m["gw"], eta=m["eta"], z=g["z"][k])
        assert rel_err(out["grad"], g["grad"][k]) < 1e-5
        assert rel_err(out["x_next"], g["x_next"][k]) < 1e-5
