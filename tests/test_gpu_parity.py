"""GPU parity tests proper: the CUDA path, called through the C ABI (ctypes),
against (a) recordings of the unmodified reference (tests/golden) and (b) the CPU
oracle on seeded inputs.  Tolerance: per-step state within 1e-5 relative in fp32
(north_star); end-to-end PSNR within 0.05 dB."""
import pytest
import torch

from oracle import dps as odps
from tests._golden import Golden, golden_names, make_network, make_problem, rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-5
DEV = "cuda:0"


@pytest.fixture(autouse=True)
def _strict_fp32():
    prev = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = prev


def _native_setup(g):
    from samplers_b200 import _native
    prob = make_problem(g, DEV)
    op = prob.operator._native_cached(torch.device(DEV))
    y = prob.operator._dense_observation(prob.observation.float())
    R = g.meta["R"]
    obs_repeat = R if len(g.meta["batch"]) and y.shape[0] > 1 else g.L
    return _native, prob, op, y, obs_repeat


@pytest.mark.parametrize("name", golden_names())
def test_k1_k2_against_reference_recording(name):
    """K1 and K2 in isolation on the reference's recorded states (teacher-forced)."""
    from samplers_b200.samplers.utils.bridge_kernels import plan_steps
    g = Golden(name)
    _native, prob, op, y, obs_repeat = _native_setup(g)
    plan = plan_steps(g["acp"], g["timesteps"].tolist(), g.meta["eta"])
    L, n = g.L, op.n
    w = prob.noise._likelihood_weight()
    worst = 0.0
    for k, sc in enumerate(plan):
        assert sc.t == g.meta["t"][k] and sc.t_prev == g.meta["t_prev"][k]
        x = g["x_t"][k].reshape(L, n).to(DEV).contiguous()
        eps = g["eps"][k].reshape(L, n).to(DEV).contiguous()
        cot = torch.empty_like(x)
        part = torch.empty(L, op.err_parts, device=DEV)
        wsb = op.workspace_bytes(L)
        ws = torch.empty(wsb // 4, device=DEV) if wsb else None
        _native.dps_pre(op, x, eps, y, obs_repeat, sc.sqrt_acp, sc.sqrt_1m_acp, w, cot, part, ws)
        assert rel_err(cot.cpu(), g["d"][k].reshape(L, n)) < TOL
        assert rel_err(part.sum(1).cpu(), g["e2"][k]) < TOL
        v = g["v"][k].reshape(L, n).to(DEV).contiguous()
        z = g["z"][k].reshape(L, n).to(DEV).contiguous()
        out, err = torch.empty_like(x), torch.empty(L, device=DEV)
        _native.dps_post(x, eps, cot, v, z if sc.std != 0 else None, part, op.err_parts, n, sc.sqrt_acp,
                         sc.sqrt_1m_acp, sc.c_ell, sc.c_s, sc.std, g.meta["gamma"], out, err)
        worst = max(worst, rel_err(out.cpu(), g["x_next"][k].reshape(L, n)))
        assert rel_err(err.cpu(), g["e2"][k].sqrt()) < TOL
    assert worst < TOL, worst


@pytest.mark.parametrize("name", golden_names())
def test_sampler_step_against_reference_recording(name):
    """The public DPSRun.step (network + K1 + VJP + K2) teacher-forced on the reference's states."""
    from samplers_b200.samplers import DPSSampler
    g = Golden(name)
    net = make_network(g, DEV)
    prob = make_problem(g, DEV)
    sampler = DPSSampler(net)
    sampler.draw = lambda shape, device, dtype: g["x_init"].to(device)
    run = sampler.prepare(prob, num_sampling_steps=g.meta["steps"], num_reconstructions=g.meta["R"],
                          gamma=g.meta["gamma"], eta=g.meta["eta"])
    try:
        assert run.num_steps == g.K
        worst = 0.0
        for k in range(g.K):
            run.x.copy_(g["x_t"][k].reshape(run.L, run.n).to(DEV))
            run.step(k, z=g["z"][k].to(DEV))
            worst = max(worst, rel_err(run.x.cpu(), g["x_next"][k].reshape(run.L, run.n)))
        assert worst < TOL, worst
    finally:
        sampler.release()


@pytest.mark.parametrize("name", ["identity_mock_eta05", "identity_ddim", "box4_gauss"])
def test_full_sampler_against_reference_output(name):
    """Free-running DPSSampler.__call__ with the reference's noise draws: final estimate + PSNR."""
    from samplers_b200.samplers import DPSSampler
    g = Golden(name)
    net, prob = make_network(g, DEV), make_problem(g, DEV)
    draws = iter([g["x_init"]] + [g["z"][k] for k in range(g.K)])
    sampler = DPSSampler(net)
    sampler.draw = lambda shape, device, dtype: next(draws).to(device)
    out = sampler(prob, num_sampling_steps=g.meta["steps"], num_reconstructions=g.meta["R"],
                  gamma=g.meta["gamma"], eta=g.meta["eta"], keep_reconstruction_dim=True).cpu()
    ref = g["x0_final"]
    assert out.shape == ref.shape
    assert rel_err(out, ref) < 1e-3
    xt = g["x_true"].unsqueeze(len(g.meta["batch"])).expand_as(ref)
    assert abs(odps.psnr(out, xt) - odps.psnr(ref, xt)) < 0.05
    assert not net.are_sampling_parameters_initialized


def test_output_shapes_and_squeeze():
    from samplers_b200.samplers import DPSSampler
    g = Golden("identity_batch")
    net, prob = make_network(g, DEV), make_problem(g, DEV)
    s = DPSSampler(net)
    assert s(prob, num_sampling_steps=4).shape == (3, *g.shape)
    assert s(prob, num_sampling_steps=4, keep_reconstruction_dim=True).shape == (3, 1, *g.shape)
    # batch > 1 AND reconstructions > 1: broken in the reference, tiled here
    assert s(prob, num_sampling_steps=4, num_reconstructions=2).shape == (3, 2, *g.shape)


def test_batch_times_reconstructions_uses_the_right_observation():
    """Sample l must see observation l // R: equal noise => reconstructions of one observation agree
    (up to cuDNN's batch-position-dependent rounding), different observations do not."""
    from samplers_b200.samplers import DPSSampler
    g = Golden("identity_batch")
    net, prob = make_network(g, DEV), make_problem(g, DEV)
    s = DPSSampler(net)
    gen = torch.Generator().manual_seed(5)
    base = torch.randn(3, 1, *g.shape, generator=gen)

    def draw(shape, device, dtype):  # identical noise for both reconstructions of each observation
        return base.expand(3, 2, *g.shape).reshape(shape).to(device)

    s.draw = draw
    out = s(prob, num_sampling_steps=4, num_reconstructions=2, gamma=0.05).cpu()
    assert rel_err(out[:, 0], out[:, 1]) < 1e-4
    assert rel_err(out[0, 0], out[1, 0]) > 1e-2


def test_kernels_are_bitwise_deterministic():
    """K1 + K2 twice on the same inputs: identical bits (fixed-order partial sums, no atomics)."""
    from samplers_b200.samplers.utils.bridge_kernels import plan_steps
    for name in ("blur9_gauss", "inpaint_gauss", "box4_gauss", "motion9_gauss"):
        g = Golden(name)
        _native, prob, op, y, obs_repeat = _native_setup(g)
        sc = plan_steps(g["acp"], g["timesteps"].tolist(), g.meta["eta"])[0]
        L, n = g.L, op.n
        x = g["x_t"][0].reshape(L, n).to(DEV).contiguous()
        eps = g["eps"][0].reshape(L, n).to(DEV).contiguous()
        v = g["v"][0].reshape(L, n).to(DEV).contiguous()
        z = g["z"][0].reshape(L, n).to(DEV).contiguous()
        outs = []
        for _ in range(2):
            cot, part = torch.empty_like(x), torch.empty(L, op.err_parts, device=DEV)
            wsb = op.workspace_bytes(L)
            ws = torch.empty(wsb // 4, device=DEV) if wsb else None
            _native.dps_pre(op, x, eps, y, obs_repeat, sc.sqrt_acp, sc.sqrt_1m_acp, 400.0, cot, part, ws)
            out = torch.empty_like(x)
            _native.dps_post(x, eps, cot, v, z, part, op.err_parts, n, sc.sqrt_acp, sc.sqrt_1m_acp, sc.c_ell,
                             sc.c_s, sc.std, 1.0, out, None)
            outs.append((cot.clone(), part.clone(), out.clone()))
        for a, b in zip(*outs):
            assert torch.equal(a, b), name


def test_tweedie_final_with_moments():
    from samplers_b200 import _native
    gen = torch.Generator().manual_seed(0)
    L, n = 5, 3 * 20 * 12
    x, e = torch.randn(L, n, generator=gen), torch.randn(L, n, generator=gen)
    acp = torch.tensor(0.37)
    sa, s1 = float(acp ** 0.5), float((1 - acp) ** 0.5)
    ref = odps.tweedie_x0(x, e, acp)
    out, tot, tsq = torch.empty(L, n, device=DEV), torch.empty(n, device=DEV), torch.empty(n, device=DEV)
    _native.tweedie(x.to(DEV), e.to(DEV), sa, s1, out, tot, tsq)
    assert torch.equal(out.cpu(), ref)  # bit-exact: same three roundings
    assert torch.allclose(tot.cpu(), ref.sum(0), rtol=1e-5, atol=1e-5)
    assert torch.allclose(tsq.cpu(), ref.square().sum(0), rtol=1e-5, atol=1e-5)
    out2 = torch.empty(L, n, device=DEV)
    _native.tweedie(x.to(DEV), e.to(DEV), sa, s1, out2)
    assert torch.equal(out2.cpu(), ref)


def test_predict_x0_matches_reference_formula_and_backward():
    g = Golden("identity_gauss")
    net = make_network(g, DEV)
    net.set_sampling_parameters(8)
    x = g["x_t"][0].to(DEV).requires_grad_()
    t = g.meta["t"][0]
    x0 = net.predict_x0(x, t)
    ref = odps.tweedie_x0(g["x_t"][0], g["eps"][0], g["acp"][t])
    assert rel_err(x0.detach().cpu(), ref) < 1e-6
    x0.square().sum().backward()
    xc = g["x_t"][0].clone().requires_grad_()
    odps.tweedie_x0(xc, g.net()(xc, t), g["acp"][t]).square().sum().backward()
    assert rel_err(x.grad.cpu(), xc.grad) < 1e-5


def test_errors_mirror_reference_conventions():
    from samplers_b200 import _native
    from samplers_b200.noise import GaussianNoise, PoissonNoise
    from samplers_b200.operators import IdentityOperator, InpaintingOperator
    with pytest.raises(ValueError):
        GaussianNoise(sigma=0.0)
    with pytest.raises(ValueError):
        PoissonNoise(rate=-1.0)
    with pytest.raises(ValueError):
        GaussianNoise(sigma=torch.ones(2))
    with pytest.raises(ValueError):
        InpaintingOperator((3, 4, 4), torch.zeros(3, 4, 5, dtype=torch.bool))
    op = IdentityOperator((3, 4, 4))._native_cached(torch.device(DEV))
    x = torch.zeros(1, 48, device=DEV)
    with pytest.raises(ValueError):  # sqrt_acp must be positive
        _native.dps_pre(op, x, x, x, 1, 0.0, 1.0, 1.0, torch.empty_like(x), torch.empty(1, op.err_parts, device=DEV), None)
    with pytest.raises(RuntimeError):
        _native.require_cuda(torch.zeros(3), "x")
