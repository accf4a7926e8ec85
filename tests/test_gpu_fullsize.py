"""BASELINE.json's full sizes, through the C ABI: one whole DPS / PGDM / PSLD / ReSample building block per
config, checked against the oracle's arithmetic executed on the device (the oracle code is the checker; it is
fp32 torch, so it runs on CUDA tensors unchanged) and against size-independent properties."""
import pytest
import torch

from oracle import dps as odps
from oracle import operators as oops
from oracle import resample as ors
from oracle.tiny_net import TinyEpsNet
from tests._golden import rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
FULL = (3, 256, 256)


@pytest.fixture(autouse=True)
def _strict_fp32():
    prev = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = prev


def _network(acp, ts):
    from samplers_b200.networks.base import EpsilonNetwork

    class Net(EpsilonNetwork):
        def __init__(self):
            super().__init__(alphas_cumprod=acp)
            self.core = TinyEpsNet(channels=3)

        def forward(self, x, t):
            return self.core(x, int(t))

        @classmethod
        def from_pretrained(cls, *a, **k):
            raise NotImplementedError

        def set_sampling_parameters(self, num_sampling_steps, batch_size=1, num_reconstructions=1):
            self._batch_size = batch_size
            self.register_buffer("timesteps", ts.to(self.alphas_cumprod.device))

        @property
        def is_condition_initialized(self):
            return True

    return Net().to(DEV)


def _dev_op(o):
    """Move an oracle operator's tensors to the device so that its torch code runs there."""
    for name in ("taps_h", "taps_v", "kernel2d", "mask", "kept"):
        if hasattr(o, name):
            setattr(o, name, getattr(o, name).to(DEV))
    return o


def _coarse_keep(shape, f, frac, seed):
    """The keep array MaskedBoxDownsampleOperator(missing_fraction=frac, seed=seed) draws."""
    g = torch.Generator().manual_seed(seed)
    return (~(torch.rand((shape[0], shape[1] // f, shape[2] // f), generator=g) < frac)).float()


CASES = {
    # config 2: Gaussian blur 61x61 sigma 3, batch 16
    "cfg2_blur_L16": (16, lambda p: p.GaussianBlurOperator(FULL), lambda: oops.OracleGaussianBlur(FULL, 61, 3.0)),
    # config 3: random mask 70 % (dense form) and 4x box super-resolution, batch 64
    "cfg3_mask_L64": (64, lambda p: p.RandomInpaintingOperator(FULL, 0.7, seed=0, flatten=False), None),
    "cfg3_box4_L64": (64, lambda p: p.BoxDownsampleOperator(FULL, 4), lambda: oops.OracleBoxDownsample(FULL, 4)),
    # config 3 read as ONE composed operator: 70 % of the 4x box-downsampled pixels dropped
    "cfg3_maskbox4_L64": (64, lambda p: p.MaskedBoxDownsampleOperator(FULL, 4, missing_fraction=0.7, seed=0),
                          lambda: oops.OracleMaskedBox(FULL, 4, _coarse_keep(FULL, 4, 0.7, 0))),
}


@pytest.mark.parametrize("name", list(CASES))
def test_full_size_dps_step_matches_oracle_on_device(name):
    """One guided DPS timestep of the public API at the config's batch size vs oracle dps_step_autograd."""
    from samplers_b200 import operators as pops
    from samplers_b200.inverse_problem import InverseProblem
    from samplers_b200.noise import GaussianNoise
    from samplers_b200.samplers import DPSSampler
    from oracle.schedule import ddpm_linear_alphas_cumprod, leading_timesteps_ascending, padded_clipped_acp
    L, make_op, make_oracle = CASES[name]
    acp, ts = padded_clipped_acp(ddpm_linear_alphas_cumprod()), leading_timesteps_ascending(1000)
    net = _network(acp, ts)
    op = make_op(pops).to(DEV)
    if make_oracle is None:  # dense mask
        keep = (~op.mask).float().to(DEV)

        class Dense(oops.OracleOperator):
            x_shape = y_shape = FULL

            def apply(self, x):
                return x * keep

        ora = Dense()
    else:
        ora = _dev_op(make_oracle())
    gen = torch.Generator(device=DEV).manual_seed(0)
    x_true = torch.rand(FULL, device=DEV, generator=gen) * 2 - 1
    y = ora.apply(x_true[None])[0] + 0.05 * torch.randn(ora.y_shape, device=DEV, generator=gen)
    prob = InverseProblem(operator=op, observation=y, noise=GaussianNoise(sigma=0.05))
    x0 = torch.randn(L, *FULL, device=DEV, generator=gen)
    z = torch.randn(L, *FULL, device=DEV, generator=gen)
    s = DPSSampler(net)
    s.draw = lambda shape, device, dtype: x0.clone()
    run = s.prepare(prob, num_sampling_steps=1000, num_reconstructions=L, gamma=1.0, eta=1.0)
    try:
        k = 400  # a mid-trajectory timestep
        sc = run.plan[k]
        run.step(k, z=z)
        ref = odps.dps_step_autograd(lambda x, t: net.core(x, int(t)), x0, t=sc.t, t_prev=sc.t_prev, s=run.timesteps[0],
                                     acp=acp.to(DEV), op=ora, y=y[None], noise_kind="gaussian",
                                     noise_param=torch.tensor(0.05, device=DEV), gamma=1.0, eta=1.0, z=z)
        assert rel_err(run.x.view(L, *FULL).cpu(), ref["x_next"].cpu()) < 1e-5
        assert rel_err(run.err.cpu(), ref["err"].cpu()) < 1e-5
    finally:
        s.release()


def test_full_size_psld_data_term_config4_shapes():
    """PSLD pixel-space block at 3x512x512 (config 4), Gaussian blur: fused node vs oracle expression on device."""
    from samplers_b200 import operators as pops
    from samplers_b200.samplers.psld import _PsldDataTerm
    shape, L = (3, 512, 512), 2
    op = pops.GaussianBlurOperator(shape).to(DEV)
    ora = _dev_op(oops.OracleGaussianBlur(shape, 61, 3.0))
    nat = op._native_cached(torch.device(DEV))
    gen = torch.Generator(device=DEV).manual_seed(1)
    x0 = torch.randn(L, *shape, device=DEV, generator=gen)
    y = torch.randn(L, *shape, device=DEV, generator=gen)
    c = torch.randn(L, *shape, device=DEV, generator=gen)
    xo = x0.clone().requires_grad_()
    hx = ora.apply(xo)
    lik_o = torch.norm(y - hx)
    xeff_o = ora.adjoint(y) + xo - ora.adjoint(hx)
    (go,) = torch.autograd.grad(0.1 * lik_o + (xeff_o * c).sum(), xo)
    xd = x0.reshape(L, -1).clone().requires_grad_()
    ws = torch.empty(nat.workspace_bytes(L) // 4, device=DEV)
    lik, xeff = _PsldDataTerm.apply(xd, nat, y.reshape(L, -1).contiguous(), 1, ws, torch.zeros(1, nat.n_y, device=DEV))
    (gd,) = torch.autograd.grad(0.1 * lik + (xeff * c.reshape(L, -1)).sum(), xd)
    assert abs(float(lik.detach()) - float(lik_o.detach())) < 1e-5 * float(lik_o.detach())
    assert rel_err(xeff.detach().cpu(), xeff_o.detach().reshape(L, -1).cpu()) < 2e-6
    assert rel_err(gd.cpu(), go.reshape(L, -1).cpu()) < 1e-5


def test_full_size_resample_latent_kernels_config5_shapes():
    """eps-DDIM step + stochastic resample on 32 x (4, 64, 64) latents (config 5: 32 samples per GPU)."""
    from samplers_b200 import _native
    from samplers_b200.samplers.resample import ddim_eps_scalars, resample_scalars
    from oracle.schedule import ddpm_linear_alphas_cumprod, padded_clipped_acp
    acp = padded_clipped_acp(ddpm_linear_alphas_cumprod())
    gen = torch.Generator(device=DEV).manual_seed(2)
    x, e, z = (torch.randn(32, 4, 64, 64, device=DEV, generator=gen) for _ in range(3))
    t, tp = 620, 610
    ref_prev, _, ref_pseudo = ors.ddim_eps_step(x, e, acp.to(DEV), t, tp, 1.0, z)
    prev, pseudo = torch.empty_like(x), torch.empty_like(x)
    _native.ddim_eps_step(x, e, z, ddim_eps_scalars(acp, t, tp, 1.0), prev, None, pseudo)
    assert rel_err(prev.cpu(), ref_prev.cpu()) < 1e-6 and rel_err(pseudo.cpu(), ref_pseudo.cpu()) < 1e-6
    sigma = ors.compute_sigma(40.0, acp[t], acp[tp])
    ref = ors.stochastic_resample(x, e, acp[tp].to(DEV), sigma.to(DEV), z)
    out = torch.empty_like(x)
    _native.stochastic_resample(x, e, z, *resample_scalars(acp, t, tp, 40.0), out)
    assert rel_err(out.cpu(), ref.cpu()) < 1e-6


def test_posterior_moments_of_many_samples_match_torch():
    """Final Tweedie + per-pixel sum / sum of squares over 64 samples at 3x256x256 (the all-reduce send buffers)."""
    from samplers_b200 import _native
    gen = torch.Generator(device=DEV).manual_seed(3)
    L, n = 64, 3 * 256 * 256
    x, e = torch.randn(L, n, device=DEV, generator=gen), torch.randn(L, n, device=DEV, generator=gen)
    acp = torch.tensor(0.42)
    out, tot, tsq = torch.empty(L, n, device=DEV), torch.empty(n, device=DEV), torch.empty(n, device=DEV)
    _native.tweedie(x, e, float(acp ** 0.5), float((1 - acp) ** 0.5), out, tot, tsq)
    ref = odps.tweedie_x0(x, e, acp.to(DEV))
    assert torch.equal(out, ref)
    assert rel_err(tot.cpu(), ref.double().sum(0).float().cpu()) < 1e-6
    assert rel_err(tsq.cpu(), ref.double().square().sum(0).float().cpu()) < 1e-6


@pytest.mark.parametrize("L,obs_repeat", [(64, 64), (64, 1), (30, 1), (50, 5)])
@pytest.mark.parametrize("path", ["tc", "tc_persist", "cuda_core_1", "cuda_core_2"])
def test_blur_k1_large_batches_every_path(L, obs_repeat, path, monkeypatch):
    """K1 of config 2's blur at batches with MORE planes than resident CTA groups -- the tensor-core kernel in several
    waves and as one wave of persistent cluster pairs (74 pairs; 90 / 150 / 192 planes) and the persistent strip kernels of the CUDA-core
    path (whose in-place interleaved h2 once raced with unfetched strips from L = 32 on) -- against the oracle's
    operator evaluated in fp64 on the device, per sample, twice (repeatability)."""
    from samplers_b200 import _native, operators as pops
    env = {"tc": {}, "tc_persist": {"PSX_TC_PERSIST": "1"}, "cuda_core_1": {"PSX_NO_TC": "1", "PSX_SPLIT": "1"},
           "cuda_core_2": {"PSX_NO_TC": "1", "PSX_SPLIT": "2"}}[path]
    for k in ("PSX_TC_PERSIST", "PSX_NO_TC", "PSX_SPLIT"):
        monkeypatch.delenv(k, raising=False)
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    _native.reload_env()
    try:
        op = pops.GaussianBlurOperator(FULL).to(DEV)
        nat = op._native_cached(torch.device(DEV))
        ora = _dev_op(oops.OracleGaussianBlur(FULL, 61, 3.0))
        ora.taps_h, ora.taps_v = ora.taps_h.double(), ora.taps_v.double()
        gen = torch.Generator(device=DEV).manual_seed(L + obs_repeat)
        n = nat.n
        x = torch.randn(L, n, device=DEV, generator=gen)
        eps = torch.randn(L, n, device=DEV, generator=gen)
        y = torch.randn(L // obs_repeat, n, device=DEV, generator=gen)
        sa, s1, w = 0.9, 0.43, 25.0
        x0 = ((x.double() - s1 * eps.double()) / sa).view(L, *FULL).requires_grad_()
        r = y.double().repeat_interleave(obs_repeat, 0).view(L, *FULL) - ora.apply(x0)
        err_ref = (r.detach() ** 2).flatten(1).sum(1)
        (g,) = torch.autograd.grad(-0.5 * (r ** 2).sum(), x0)          # = A^T r
        cot_ref = (g * (w / sa)).flatten(1)
        ws = torch.empty(nat.workspace_bytes(L) // 4, device=DEV)
        outs = []
        for _ in range(2):
            cot = torch.full_like(x, float("nan"))
            part = torch.full((L, nat.err_parts), float("nan"), device=DEV)
            _native.dps_pre(nat, x, eps, y, obs_repeat, sa, s1, w, cot, part, ws)
            torch.cuda.synchronize()
            per = ((cot.double() - cot_ref).norm(dim=1) / cot_ref.norm(dim=1))
            assert float(per.max()) < 1e-5, [int(i) for i in torch.nonzero(per > 1e-5).flatten()]
            assert float(((part.double().sum(1) - err_ref).abs() / err_ref).max()) < 1e-5
            outs.append(cot)
        assert torch.equal(outs[0], outs[1])
    finally:
        for k in env:
            monkeypatch.delenv(k, raising=False)
        _native.reload_env()
