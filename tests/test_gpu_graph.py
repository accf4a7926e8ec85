"""CUDA-graph replay of the DPS / PGDM timestep: the *_dev entry points, which read the step scalars
from a device row, must be bit-identical to the by-value ones, and a graphed run must reproduce the eager run."""
import pytest
import torch

from tests._golden import rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
SHAPE = (3, 32, 32)


def _ops():
    from samplers_b200 import operators as P
    return {
        "identity": P.IdentityOperator(SHAPE),
        "mask": P.RandomInpaintingOperator(SHAPE, 0.7, seed=0, flatten=False),
        "box": P.BoxDownsampleOperator(SHAPE, 4),
        "blur": P.GaussianBlurOperator(SHAPE, kernel_size=9, sigma=1.5),
        "motion": P.MotionBlurOperator(SHAPE, kernel_size=9, angle_deg=30.0),
    }


@pytest.mark.parametrize("name", ["identity", "mask", "box", "blur", "motion"])
def test_dev_entry_points_are_bitwise_equal_to_by_value(name):
    from samplers_b200 import _native
    op = _ops()[name].to(DEV)
    nat = op._native_cached(torch.device(DEV))
    L, n = 5, nat.n
    gen = torch.Generator(device=DEV).manual_seed(0)
    x, e, v, z = (torch.randn(L, n, device=DEV, generator=gen) for _ in range(4))
    y = torch.randn(1, nat.n_y, device=DEV, generator=gen)
    if name == "mask":
        y = y * (~op.mask).float().reshape(1, -1).to(DEV)
    sa, s1, w = 0.8366600275039673, 0.547722578048706, 400.0
    c_ell, c_s, sd, gamma = 0.97, 0.021, 0.11, 1.3
    row = torch.tensor([[sa, s1, float(torch.tensor(w) / torch.tensor(sa)), c_ell, c_s, sd, gamma, 0.0]], device=DEV)
    wsb = nat.workspace_bytes(L)
    ws = torch.empty(wsb // 4, device=DEV) if wsb else None

    def run(dev_row):
        cot, part = torch.empty(L, n, device=DEV), torch.empty(L, nat.err_parts, device=DEV)
        out, err = torch.empty(L, n, device=DEV), torch.empty(L, device=DEV)
        if dev_row:
            _native.dps_pre_dev(nat, x, e, y, L, row, cot, part, ws)
            _native.dps_post_dev(x, e, cot, v, z, part, nat.err_parts, n, row, out, err)
        else:
            _native.dps_pre(nat, x, e, y, L, sa, s1, w, cot, part, ws)
            _native.dps_post(x, e, cot, v, z, part, nat.err_parts, n, sa, s1, c_ell, c_s, sd, gamma, out, err)
        return cot, out, err

    a, b = run(False), run(True)
    for p, q in zip(a, b):
        assert torch.equal(p, q)
    # in place (x_next aliases x_t), as the graphed step uses it
    xa = x.clone()
    cot, part = torch.empty(L, n, device=DEV), torch.empty(L, nat.err_parts, device=DEV)
    _native.dps_pre_dev(nat, xa, e, y, L, row, cot, part, ws)
    _native.dps_post_dev(xa, e, cot, v, z, part, nat.err_parts, n, row, xa, None)
    assert torch.equal(xa, a[1])


def _problem(op_name, sigma=0.05):
    from samplers_b200.inverse_problem import InverseProblem
    from samplers_b200.noise import GaussianNoise
    op = _ops()[op_name].to(DEV)
    gen = torch.Generator(device=DEV).manual_seed(1)
    x_true = torch.rand(SHAPE, device=DEV, generator=gen) * 2 - 1
    y = op.apply(x_true[None])[0]
    y = y + sigma * torch.randn(y.shape, device=DEV, generator=gen)
    if op_name == "mask":
        y = y * (~op.mask).float().to(DEV)
    return InverseProblem(operator=op, observation=y, noise=GaussianNoise(sigma=sigma))


def _tiny_net():
    from samplers_b200.networks import DDPMNetwork
    torch.manual_seed(1234)
    return DDPMNetwork.from_config("tiny", device=DEV)


class _Tape:
    """Recorded N(0,1) draws handed out in order (the sampler's ``draw`` hook)."""

    def __init__(self, shape, count):
        gen = torch.Generator(device=DEV).manual_seed(7)
        self.items = [torch.randn(shape, device=DEV, generator=gen) for _ in range(count)]
        self.i = 0

    def __call__(self, shape, device, dtype):
        t = self.items[self.i]
        self.i += 1
        return t.clone()


@pytest.mark.parametrize("sampler_name,op_name", [("dps", "blur"), ("dps", "mask"), ("dps", "box"), ("pgdm", "box")])
def test_graphed_run_reproduces_eager_run(sampler_name, op_name):
    from samplers_b200.samplers import DPSSampler, PGDMSampler
    net = _tiny_net()
    prob = _problem(op_name)
    cls = DPSSampler if sampler_name == "dps" else PGDMSampler
    outs = []
    for graph in (False, True):
        s = cls(net, cuda_graph=graph)
        s.draw = _Tape((4, *SHAPE), 16)
        outs.append(s(prob, num_sampling_steps=12, num_reconstructions=4))
    assert outs[0].shape == outs[1].shape == (4, *SHAPE)
    assert torch.isfinite(outs[1]).all()
    # one step: K1 bitwise, x to 2e-4 (test below); over a whole run of this random-init network the state grows to ~1e4
    # and last-bit differences of cuDNN / cuBLAS algorithm choices under capture are amplified
    assert rel_err(outs[1].cpu(), outs[0].cpu()) < 5e-4


def test_graph_replay_out_of_order_and_restart():
    """step(k) with a k that is not the device counter's value re-seeks the table; replaying a step twice from the
    same state gives the same result."""
    from samplers_b200.samplers import DPSSampler
    net = _tiny_net()
    prob = _problem("blur")
    s = DPSSampler(net, cuda_graph=True)
    tape = _Tape((2, *SHAPE), 4)
    s.draw = tape
    run = s.prepare(prob, num_sampling_steps=10, num_reconstructions=2)
    try:
        run.capture()
        x0 = run.x.clone()
        z = tape.items[1].reshape(2, -1)
        run.step(5, z=z)
        a, cot_a, err_a = run.x.clone(), run.cot.clone(), run.err.clone()
        run.x.copy_(x0)
        run.step(5, z=z)
        # K1 is bitwise reproducible; x goes through the network VJP, whose cuDNN backward kernels use atomics
        # (eager repeats of this step differ by ~2e-5 as well)
        assert torch.equal(run.cot, cot_a) and torch.equal(run.err, err_a)
        assert rel_err(run.x.cpu(), a.cpu()) < 2e-4
        assert int(run.k_dev) == 6
        # eager single step from the same state with the same scalars
        eager = DPSSampler(net)
        eager.draw = lambda *_: x0.view(2, *SHAPE).clone()
        r2 = eager.prepare(prob, num_sampling_steps=10, num_reconstructions=2)
        r2.step(5, z=z)
        assert torch.equal(r2.cot, cot_a) and torch.equal(r2.err, err_a)
        assert rel_err(r2.x.cpu(), a.cpu()) < 2e-4
    finally:
        s.release()


def test_graph_draws_fresh_noise_each_replay():
    from samplers_b200.samplers import DPSSampler
    net = _tiny_net()
    prob = _problem("identity")
    s = DPSSampler(net, cuda_graph=True)
    run = s.prepare(prob, num_sampling_steps=10, num_reconstructions=2)
    try:
        run.capture()
        assert run._draw_in_graph
        run.step(0)
        z0 = run.z.clone()
        run.step(1)
        assert not torch.equal(z0, run.z)
        assert abs(float(run.z.mean())) < 0.1 and abs(float(run.z.std()) - 1.0) < 0.1
        assert torch.isfinite(run.x).all()
        with pytest.raises(ValueError):
            run.step(2, z=z0)                      # a graph that draws its own noise cannot take injected noise
    finally:
        s.release()


def test_network_that_syncs_on_the_timestep_fails_loudly_under_capture():
    from samplers_b200.samplers import DPSSampler
    from tests.test_gpu_fullsize import _network
    from oracle.schedule import ddpm_linear_alphas_cumprod, leading_timesteps_ascending, padded_clipped_acp
    acp, ts = padded_clipped_acp(ddpm_linear_alphas_cumprod()), leading_timesteps_ascending(10)
    net = _network(acp, ts)          # forward() does int(t): a device->host sync
    s = DPSSampler(net, cuda_graph=True)
    with pytest.raises(Exception):
        s(_problem("identity"), num_sampling_steps=10, num_reconstructions=1)
    torch.cuda.synchronize()
    # the device is still usable afterwards, eagerly
    out = DPSSampler(net)(_problem("identity"), num_sampling_steps=10, num_reconstructions=1)
    assert torch.isfinite(out).all()


@pytest.mark.parametrize("graph", [False, True])
def test_half_precision_network_with_fp32_state(graph):
    """A bf16 network (the reference's latent pipelines run in bf16) on the fp32 state: same trajectory as the fp32
    network up to the network's own rounding; eager and graph."""
    from samplers_b200.networks import DDPMNetwork
    from samplers_b200.samplers import DPSSampler
    prob = _problem("blur")
    outs = {}
    for dt in (torch.float32, torch.bfloat16):
        torch.manual_seed(1234)
        net = DDPMNetwork.from_config("tiny", device=DEV, torch_dtype=dt)
        s = DPSSampler(net, cuda_graph=graph, philox_seed=3)
        x_init = torch.randn(2, *SHAPE, device=DEV, generator=torch.Generator(device=DEV).manual_seed(5))
        s.draw = lambda shape, device, dtype: x_init.clone()
        run = s.prepare(prob, num_sampling_steps=6, num_reconstructions=2, gamma=0.05)
        try:
            if graph:
                run.capture()
            run.step(0)
            outs[dt] = run.x.clone()
            assert run.x.dtype == torch.float32 and run.cot.dtype == torch.float32
            assert torch.isfinite(run.finalize()).all()
        finally:
            s.release()
    assert rel_err(outs[torch.bfloat16].cpu(), outs[torch.float32].cpu()) < 5e-2
