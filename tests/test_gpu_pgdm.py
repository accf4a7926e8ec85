"""PGDM on the CUDA path vs recordings of the unmodified reference PGDMSampler."""
import pytest
import torch

from tests._golden import PgdmGolden, make_network, make_pgdm_problem, pgdm_names, rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(autouse=True)
def _strict_fp32():
    prev = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = prev


@pytest.mark.parametrize("name", pgdm_names())
def test_pgdm_steps_teacher_forced(name):
    from samplers_b200.samplers.dps import DPSRun
    from samplers_b200.samplers.utils import BatchView
    g = PgdmGolden(name)
    m = g.meta
    net, prob = make_network(g, DEV), make_pgdm_problem(g, DEV)
    net.set_sampling_parameters(m["steps"])
    view = BatchView(prob.batch_shape, m["R"], prob.operator.x_shape)
    gw = torch.tensor(m["gw"], dtype=torch.float32)
    run = DPSRun(net, prob, view, 0.0, m["eta"], lambda s, d, t: g["x_init"].to(d),
                 weight=2.0 * prob.operator._pinv_gain(),
                 fixed_scale=lambda sc: float(gw * torch.tensor(sc.sqrt_1m_acp, dtype=torch.float32)))
    worst = 0.0
    for k in range(g.K):
        run.x.copy_(g["x_t"][k].reshape(run.L, run.n).to(DEV))
        run.step(k, z=g["z"][k].to(DEV))
        worst = max(worst, rel_err(run.x.cpu(), g["x_next"][k].reshape(run.L, run.n)))
    assert worst < 1e-5, worst


@pytest.mark.parametrize("name", pgdm_names())
def test_pgdm_full_run(name):
    from samplers_b200.samplers import PGDMSampler
    g = PgdmGolden(name)
    m = g.meta
    net, prob = make_network(g, DEV), make_pgdm_problem(g, DEV)
    draws = iter([g["x_init"]] + [g["z"][k] for k in range(g.K)])
    s = PGDMSampler(net)
    s.draw = lambda shape, device, dtype: next(draws).to(device)
    out = s(prob, num_sampling_steps=m["steps"], num_reconstructions=m["R"], guidance_weight=m["gw"], eta=m["eta"],
            keep_reconstruction_dim=True).cpu()
    assert out.shape == g["x0_final"].shape
    assert rel_err(out, g["x0_final"]) < 1e-3
    assert not net.are_sampling_parameters_initialized


def test_pgdm_shapes_and_missing_pseudo_inverse():
    """tests/samplers/test_pgdm.py:107-124 of the reference + the NotImplementedError contract (pgdm.py:55-66)."""
    from samplers_b200.operators import GaussianBlurOperator
    from samplers_b200.inverse_problem import InverseProblem
    from samplers_b200.noise import GaussianNoise
    from samplers_b200.samplers import PGDMSampler
    g = PgdmGolden("identity")
    net, prob = make_network(g, DEV), make_pgdm_problem(g, DEV)
    out = PGDMSampler(net)(prob, num_sampling_steps=5, num_reconstructions=2, guidance_weight=0.1)
    assert out.shape == (2, 2, *g.shape)
    single = InverseProblem(operator=prob.operator, observation=prob.observation[0], noise=prob.noise)
    assert PGDMSampler(net)(single, num_sampling_steps=5, guidance_weight=0.1).shape == g.shape
    blur = InverseProblem(operator=GaussianBlurOperator(g.shape, 9, 1.5).to(DEV), observation=prob.observation[0],
                          noise=GaussianNoise(sigma=0.05))
    with pytest.raises(NotImplementedError, match="apply_pseudo_inverse"):
        PGDMSampler(net)(blur, num_sampling_steps=5)
