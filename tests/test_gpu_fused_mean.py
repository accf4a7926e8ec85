"""psx_dps_pre_mean / psx_dps_post_mean (the tensor-core blur's K1 also writes the bridge mean, K2 reads it in place
of x_t and eps), through the C ABI: the pair must be BIT-IDENTICAL to psx_dps_pre + psx_dps_post, the mean itself must
be the separately rounded torch expression of bridge_kernels.py:41 over Tweedie's x0 (networks/base.py:42-43), and the
sampler must produce the same state with and without it."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
FULL = (3, 256, 256)
N = 3 * 256 * 256


def _blur_op():
    from samplers_b200 import operators as pops
    return pops.GaussianBlurOperator(FULL).to(DEV)


def _inputs(L, seed=0):
    g = torch.Generator(device=DEV).manual_seed(seed)
    x = torch.randn(L, N, device=DEV, generator=g)
    eps = torch.randn(L, N, device=DEV, generator=g)
    y = torch.rand(1, N, device=DEV, generator=g) * 2 - 1
    v = torch.randn(L, N, device=DEV, generator=g)
    z = torch.randn(L, N, device=DEV, generator=g)
    return x, eps, y, v, z


SCALARS = [  # (sa, s1, c_ell, c_s, std): a late, a middle and the first (noisiest) timestep
    (0.9969, 0.0787, 0.31, 0.69, 0.021),
    (0.6, 0.8, 0.55, 0.4, 0.3),
    (0.0063, 0.99998, 0.97, 0.004, 0.05),
]


@pytest.mark.parametrize("L,persist", [(1, False), (2, False), (16, False), (16, True)])
@pytest.mark.parametrize("sc", SCALARS)
def test_pair_with_mean_is_bit_identical(L, persist, sc):
    from samplers_b200 import _native
    sa, s1, c_ell, c_s, std = sc
    weight, gamma = 400.0, 1.0
    if persist:
        os.environ["PSX_TC_PERSIST"] = "1"
        _native.load().psx_reload_env()
    try:
        op = _blur_op()._native_cached(torch.device(DEV))
        assert op.fuses_mean(L)
        x, eps, y, v, z = _inputs(L)
        ws = torch.empty(max(op.workspace_bytes(L) // 4, 1), device=DEV)
        parts = op.err_parts
        cot_a, cot_b = torch.empty_like(x), torch.empty_like(x)
        ep_a, ep_b = torch.empty(L, parts, device=DEV), torch.empty(L, parts, device=DEV)
        out_a, out_b, mean = torch.empty_like(x), torch.empty_like(x), torch.full_like(x, float("nan"))
        err_a, err_b = torch.empty(L, device=DEV), torch.empty(L, device=DEV)
        _native.dps_pre(op, x, eps, y, L, sa, s1, weight, cot_a, ep_a, ws)
        _native.dps_post(x, eps, cot_a, v, z, ep_a, parts, N, sa, s1, c_ell, c_s, std, gamma, out_a, err_a)
        _native.dps_pre_mean(op, x, eps, y, L, sa, s1, weight, c_ell, c_s, cot_b, ep_b, mean, ws)
        _native.dps_post_mean(mean, cot_b, v, z, ep_b, parts, N, s1, std, gamma, out_b, err_b)
        torch.cuda.synchronize()
        assert torch.equal(cot_a, cot_b) and torch.equal(ep_a, ep_b) and torch.equal(err_a, err_b)
        assert torch.isfinite(out_a).all()
        assert torch.equal(out_a, out_b)
        # the mean itself: bridge_kernels.py:41 over networks/base.py:42-43, one rounding per torch op
        f32 = lambda v: torch.tensor(v, dtype=torch.float32, device=DEV)  # noqa: E731  (tensor divisor: true division)
        x0 = (x - f32(s1) * eps) / f32(sa)
        ref = f32(c_ell) * x + f32(c_s) * x0
        assert torch.equal(mean, ref) or float((mean - ref).abs().max() / ref.abs().max()) < 2e-7
        # device-row entry points: same bits; x_next written in place over the state
        row = torch.tensor([[sa, s1, weight / sa, c_ell, c_s, std, gamma, 0.0]], device=DEV)
        row[0, 2] = torch.tensor(weight, dtype=torch.float32) / torch.tensor(sa, dtype=torch.float32)
        cot_c, ep_c, mean_c = torch.empty_like(x), torch.empty_like(ep_a), torch.empty_like(x)
        _native.dps_pre_mean(op, x, eps, y, L, 1.0, 0.0, 1.0, 0.0, 0.0, cot_c, ep_c, mean_c, ws, step_row=row)
        state = x.clone()
        _native.dps_post_mean(mean_c, cot_c, v, z, ep_c, parts, N, 0.0, 0.0, 0.0, state, None, step_row=row)
        torch.cuda.synchronize()
        assert torch.equal(mean_c, mean) and torch.equal(cot_c, cot_b) and torch.equal(state, out_b)
        # the step's noise folded into the mean by K1 (mean + std z): K2 then reads mean, cot and the VJP only
        cot_d, ep_d, mean_z, out_d = torch.empty_like(x), torch.empty_like(ep_a), torch.empty_like(x), torch.empty_like(x)
        _native.dps_pre_mean(op, x, eps, y, L, sa, s1, weight, c_ell, c_s, cot_d, ep_d, mean_z, ws, z=z, std=std)
        _native.dps_post_mean(mean_z, cot_d, v, None, ep_d, parts, N, s1, 0.0, gamma, out_d, None)
        torch.cuda.synchronize()
        assert torch.equal(mean_z, mean + f32(std) * z) and torch.equal(cot_d, cot_a) and torch.equal(out_d, out_a)
        mean_z.fill_(float("nan"))
        _native.dps_pre_mean(op, x, eps, y, L, 1.0, 0.0, 1.0, 0.0, 0.0, cot_d, ep_d, mean_z, ws, step_row=row, z=z)
        state = x.clone()
        _native.dps_post_mean(mean_z, cot_d, v, None, ep_d, parts, N, 0.0, 0.0, 0.0, state, None, step_row=row)
        torch.cuda.synchronize()
        assert torch.equal(state, out_a)
    finally:
        if persist:
            del os.environ["PSX_TC_PERSIST"]
            _native.load().psx_reload_env()


def test_no_noise_and_fixed_scale_modes():
    """std == 0 (d_z NULL) and the fixed-scale (PGDM) form of K2."""
    from samplers_b200 import _native
    sa, s1, c_ell, c_s = 0.8, 0.6, 0.5, 0.45
    L = 3
    op = _blur_op()._native_cached(torch.device(DEV))
    x, eps, y, v, _ = _inputs(L, seed=1)
    ws = torch.empty(max(op.workspace_bytes(L) // 4, 1), device=DEV)
    cot, ep, mean = torch.empty_like(x), torch.empty(L, op.err_parts, device=DEV), torch.empty_like(x)
    out_a, out_b = torch.empty_like(x), torch.empty_like(x)
    _native.dps_pre_mean(op, x, eps, y, L, sa, s1, 1.0, c_ell, c_s, cot, ep, mean, ws)
    _native.dps_post(x, eps, cot, v, None, None, 0, N, sa, s1, c_ell, c_s, 0.0, 0.37, out_a, None)
    _native.dps_post_mean(mean, cot, v, None, None, 0, N, s1, 0.0, 0.37, out_b, None)
    torch.cuda.synchronize()
    assert torch.equal(out_a, out_b)


def test_other_operators_refuse():
    from samplers_b200 import _native
    from samplers_b200 import operators as pops
    op = pops.IdentityOperator(FULL).to(DEV)._native_cached(torch.device(DEV))
    assert not op.fuses_mean(1)
    x, eps, y, _, _ = _inputs(1)
    cot, ep, mean = torch.empty_like(x), torch.empty(1, op.err_parts, device=DEV), torch.empty_like(x)
    with pytest.raises(NotImplementedError):
        _native.dps_pre_mean(op, x, eps, y, 1, 0.8, 0.6, 1.0, 0.5, 0.4, cot, ep, mean, None)
    # a blur the tensor-core kernel does not take (128 x 128 planes) refuses as well
    small = pops.GaussianBlurOperator((3, 128, 128), 9, 1.5).to(DEV)._native_cached(torch.device(DEV))
    assert not small.fuses_mean(1)
    # ... and so does config 2's blur at batches that leave no SM idle (the classic pair runs there)
    big = _blur_op()._native_cached(torch.device(DEV))
    assert big.fuses_mean(16) and not big.fuses_mean(24) and not big.fuses_mean(64)
    L = 24
    x, eps, y, _, _ = _inputs(L)
    ws = torch.empty(max(big.workspace_bytes(L) // 4, 1), device=DEV)
    cot, ep, mean = torch.empty_like(x), torch.empty(L, big.err_parts, device=DEV), torch.empty_like(x)
    with pytest.raises(NotImplementedError):
        _native.dps_pre_mean(big, x, eps, y, L, 0.8, 0.6, 1.0, 0.5, 0.4, cot, ep, mean, ws)


@pytest.mark.parametrize("graph", [False, True])
def test_sampler_state_identical_with_and_without(graph):
    """DPSRun on config 2's operator: four guided timesteps, eager and as a replayed CUDA graph, with the fused mean
    (the default for this operator) and with the classic K1 / K2 pair."""
    from oracle.schedule import ddpm_linear_alphas_cumprod, leading_timesteps_ascending, padded_clipped_acp
    from samplers_b200.inverse_problem import InverseProblem
    from samplers_b200.noise import GaussianNoise
    from samplers_b200.samplers import DPSSampler
    from samplers_b200.networks.base import EpsilonNetwork

    class Net(EpsilonNetwork):  # a two-layer conv eps-net whose forward takes the timestep as a device tensor
        def __init__(self, acp, ts):
            super().__init__(alphas_cumprod=acp)
            self._ts = ts
            self.c1 = torch.nn.Conv2d(3, 8, 3, padding=1)
            self.c2 = torch.nn.Conv2d(8, 3, 3, padding=1)

        def forward(self, x, t):
            t = torch.as_tensor(t, device=x.device).to(x.dtype).reshape(-1)[:1]
            return self.c2(torch.tanh(self.c1(x))) * torch.cos(t * 0.01).view(1, 1, 1, 1) + 0.5 * x

        @classmethod
        def from_pretrained(cls, *a, **k):
            raise NotImplementedError

        def set_sampling_parameters(self, num_sampling_steps, batch_size=1, num_reconstructions=1):
            self._batch_size = batch_size
            self.register_buffer("timesteps", self._ts.to(self.alphas_cumprod.device))

        @property
        def is_condition_initialized(self):
            return True

    prev = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.deterministic)
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.deterministic = True   # the network's backward-data convolution must not use atomics
    try:
        acp, ts = padded_clipped_acp(ddpm_linear_alphas_cumprod()), leading_timesteps_ascending(1000)
        torch.manual_seed(5)
        net = Net(acp, ts).to(DEV)
        op = _blur_op()
        g = torch.Generator(device=DEV).manual_seed(3)
        y = torch.rand(FULL, device=DEV, generator=g) * 2 - 1
        prob = InverseProblem(operator=op, observation=y, noise=GaussianNoise(sigma=0.05))
        L = 2
        x0 = torch.randn(L, *FULL, device=DEV, generator=g)
        zs = [torch.randn(L, *FULL, device=DEV, generator=g) for _ in range(4)]
        states = []
        for fused in (True, False, False):
            s = DPSSampler(net)
            s.draw = lambda shape, device, dtype: x0.clone()
            run = s.prepare(prob, num_sampling_steps=1000, num_reconstructions=L, gamma=1.0, eta=1.0)
            try:
                assert run._fused_mean
                run._fused_mean = fused
                if graph:
                    run.capture(draw_in_graph=False)
                for k in range(4):
                    run.step(k, z=zs[k])
                torch.cuda.synchronize()
                states.append(run.x.clone())
            finally:
                s.release()
        assert torch.isfinite(states[0]).all()
        if torch.equal(states[1], states[2]):   # the classic pair reproduces itself bit for bit: so must the fused one
            assert torch.equal(states[0], states[1])
        else:                                   # (a network whose backward is not run-to-run reproducible)
            assert float((states[0] - states[1]).norm() / states[1].norm()) < 1e-5
    finally:
        (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32,
         torch.backends.cudnn.deterministic) = prev
