"""PSLD on the CUDA path vs recordings of the unmodified reference PSLDSampler (tests/golden/psld_*.npz)."""
import pytest
import torch

from tests._golden import PsldGolden, make_latent_network, make_psld_problem, psld_names, rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(autouse=True)
def _strict_fp32():
    prev = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = prev


@pytest.mark.parametrize("name", psld_names())
def test_psld_full_run_matches_reference(name):
    from samplers_b200.samplers import PSLDSampler
    g = PsldGolden(name)
    m = g.meta
    net, prob = make_latent_network(g, DEV), make_psld_problem(g, DEV)
    draws = iter([g["z_init"]] + [g["noise"][k] for k in range(g.K)])
    s = PSLDSampler(net)
    s.draw = lambda shape, device, dtype: next(draws).to(device)
    out = s(prob, num_sampling_steps=m["steps"], num_reconstructions=m["R"], gamma=m["gamma"], omega=m["omega"],
            eta=m["eta"]).cpu()
    assert out.shape == g["x_out"].shape
    assert rel_err(out, g["x_out"]) < 1e-4
    assert not net.are_sampling_parameters_initialized
    lat = PSLDSampler(net)
    lat.draw = lambda shape, device, dtype: torch.zeros(shape, device=device, dtype=dtype)
    z = lat(prob, num_sampling_steps=4, decode_output=False)
    assert z.shape == (*m["batch"], 1, *m["latent_shape"])


@pytest.mark.parametrize("name", psld_names())
def test_psld_data_term_forward_backward(name):
    """The fused pixel-space block (K1-based autograd node) vs the oracle's torch expression."""
    from samplers_b200.samplers.psld import _PsldDataTerm
    g = PsldGolden(name)
    op_o = g.oracle_op()
    prob = make_psld_problem(g, DEV)
    nat = prob.operator._native_cached(torch.device(DEV))
    L = g.L
    y_flat = g.y_flat(op_o)
    y_dev = prob.operator._dense_observation(prob.observation.float())
    obs_repeat = g.meta["R"] if len(g.meta["batch"]) else L
    gen = torch.Generator().manual_seed(0)
    x0 = torch.randn(L, *g.shape, generator=gen)
    c_x = torch.randn(L, *g.shape, generator=gen)
    # oracle
    xo = x0.clone().requires_grad_()
    hx = op_o.apply(xo)
    lik_o = torch.norm(y_flat - hx)
    xeff_o = op_o.adjoint(y_flat) + xo - op_o.adjoint(hx)
    (go,) = torch.autograd.grad(0.3 * lik_o + (xeff_o * c_x).sum(), xo)
    # product
    xd = x0.reshape(L, -1).to(DEV).requires_grad_()
    wsb = nat.workspace_bytes(L)
    ws = torch.empty(wsb // 4, device=DEV) if wsb else None
    zeros_y = torch.zeros(1, nat.n_y, device=DEV)
    lik, xeff = _PsldDataTerm.apply(xd, nat, y_dev, obs_repeat, ws, zeros_y)
    (gd,) = torch.autograd.grad(0.3 * lik + (xeff * c_x.reshape(L, -1).to(DEV)).sum(), xd)
    assert abs(float(lik.detach()) - float(lik_o.detach())) < 1e-5 * float(lik_o.detach())
    assert rel_err(xeff.detach().cpu(), xeff_o.detach().reshape(L, -1)) < 2e-6
    assert rel_err(gd.cpu(), go.reshape(L, -1)) < 1e-5


def test_psld_requires_latent_network():
    from samplers_b200.samplers import PSLDSampler
    from tests._golden import Golden, make_network
    with pytest.raises(TypeError):
        PSLDSampler(make_network(Golden("identity_gauss"), DEV))


def test_bridge_update_and_lincomb_kernels():
    from oracle import dps as odps
    from samplers_b200 import _native
    gen = torch.Generator().manual_seed(1)
    x, e, z, g = (torch.randn(3, 1001, generator=gen) for _ in range(4))
    acp = torch.tensor([1.0, 0.9, 0.6, 0.3])
    c_ell, c_s, std = odps.bridge_coefficients(acp, 3, 2, 0, 0.7)
    ref = (c_ell * x + c_s * odps.tweedie_x0(x, e, acp[3])).float() + std.float() * z - g
    out = torch.empty(3, 1001, device=DEV)
    _native.bridge_update(x.to(DEV), e.to(DEV), z.to(DEV), g.to(DEV), float(acp[3] ** 0.5), float((1 - acp[3]) ** 0.5),
                          float(c_ell.float()), float(c_s.float()), float(std.float()), -1.0, out)
    assert rel_err(out.cpu(), ref) < 1e-6
    out2 = torch.empty(3, 1001, device=DEV)
    _native.lincomb3(x.to(DEV), 2.0, e.to(DEV), -0.5, z.to(DEV), 0.25, out2)
    assert torch.allclose(out2.cpu(), 2.0 * x - 0.5 * e + 0.25 * z, atol=1e-6)


def test_psld_and_resample_accept_a_bf16_network_on_fp32_state():
    """The reference runs its latent pipelines in bf16 (scripts/run_psld.py:14): network / VAE in bf16, sampler
    state and kernels fp32 -- output close to the all-fp32 run (bf16 network rounding only), fp32 dtype."""
    from samplers_b200.samplers import PSLDSampler, ReSampleSampler
    from tests._golden import ResampleGolden, make_resample_problem
    g = PsldGolden("identity")
    m = g.meta
    outs = {}
    for dt in (torch.float32, torch.bfloat16):
        net, prob = make_latent_network(g, DEV), make_psld_problem(g, DEV)
        net.core.to(dt)                                  # weights only; the schedule buffer stays fp32
        # concrete networks report their pipeline's dtype (networks/diffusers/stable_diffusion.py:353-355)
        type(net).dtype = property(lambda self: next(self.core.parameters()).dtype)
        assert net.dtype == dt
        draws = iter([g["z_init"]] + [g["noise"][k] for k in range(g.K)])
        s = PSLDSampler(net)
        s.draw = lambda shape, device, dtype: next(draws).to(device)
        outs[dt] = s(prob, num_sampling_steps=m["steps"], num_reconstructions=m["R"], gamma=m["gamma"],
                     omega=m["omega"], eta=m["eta"])
        assert outs[dt].dtype == torch.float32 and torch.isfinite(outs[dt]).all()
    assert rel_err(outs[torch.bfloat16].cpu(), outs[torch.float32].cpu()) < 0.1
    gr = ResampleGolden("identity")
    net, prob = make_latent_network(gr, DEV), make_resample_problem(gr, DEV)
    net.core.to(torch.bfloat16)
    type(net).dtype = property(lambda self: next(self.core.parameters()).dtype)
    draws = iter(gr["draws"])
    s = ReSampleSampler(net)
    s.draw = lambda shape, device, dtype: next(draws).to(device)
    try:
        out = s(prob, num_sampling_steps=gr.meta["steps"], num_reconstructions=gr.meta["R"], **gr.meta["kw"])
    except StopIteration:       # bf16 rounding may change the number of optimiser iterations / draws
        out = None
    assert out is None or (out.dtype == torch.float32 and torch.isfinite(out).all())


def test_psld_and_resample_run_on_the_stable_diffusion_adapter():
    """The SD-shaped adapter (narrow random-init config) through both latent samplers: shapes, finiteness, clean-up."""
    from samplers_b200 import operators as P
    from samplers_b200.inverse_problem import InverseProblem
    from samplers_b200.networks import StableDiffusionCondition, StableDiffusionNetwork
    from samplers_b200.noise import GaussianNoise
    from samplers_b200.samplers import PSLDSampler, ReSampleSampler
    net = StableDiffusionNetwork.from_config("sd15-tiny", device=DEV)
    shape = (3, 64, 64)
    op = P.GaussianBlurOperator(shape, 9, 1.5).to(DEV)
    x = torch.rand(shape, device=DEV, generator=torch.Generator(device=DEV).manual_seed(0)) * 2 - 1
    prob = InverseProblem.from_clean_data(x, operator=op, noise=GaussianNoise(sigma=0.05),
                                          rng=torch.Generator(device=DEV).manual_seed(1))
    cond = StableDiffusionCondition(guidance_scale=2.0, prompt_embeds=torch.zeros(1, 7, 32))
    out = PSLDSampler(net)(prob, num_sampling_steps=6, num_reconstructions=2, condition=cond)
    assert out.shape == (2, *shape) and torch.isfinite(out).all()
    lat = PSLDSampler(net)(prob, num_sampling_steps=4, decode_output=False)
    assert lat.shape == (1, 4, 8, 8)
    out = ReSampleSampler(net)(prob, num_sampling_steps=8, num_reconstructions=2, max_optimization_iters=3,
                               time_travel_interval=2, inter_timesteps=2)
    assert out.shape == (2, *shape) and torch.isfinite(out).all()
    assert not net.is_condition_initialized and not net.are_sampling_parameters_initialized
