"""ReSample on the CUDA path vs recordings of the unmodified reference ReSampleSampler and the CPU oracle."""
import pytest
import torch

from oracle import resample as ors
from tests._golden import (ResampleGolden, make_latent_network, make_resample_problem, rel_err, resample_names)

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(autouse=True)
def _strict_fp32():
    prev = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = prev


@pytest.mark.parametrize("name", resample_names())
def test_resample_full_run_matches_reference(name):
    from samplers_b200.samplers import ReSampleSampler
    g = ResampleGolden(name)
    m = g.meta
    net, prob = make_latent_network(g, DEV), make_resample_problem(g, DEV)
    draws = iter(g["draws"])
    calls = []
    real_forward = net.forward

    def spy(x, t):
        calls.append(int(t))
        return real_forward(x, t)

    net.forward = spy
    s = ReSampleSampler(net)
    s.draw = lambda shape, device, dtype: next(draws).to(device)
    out = s(prob, num_sampling_steps=m["steps"], num_reconstructions=m["R"], **m["kw"]).cpu()
    assert calls == g["call_t"].tolist()                     # same time-travel / stage control flow
    assert next(draws, None) is None                         # consumed exactly the reference's noise draws
    assert out.shape == g["x_out"].shape
    assert rel_err(out, g["x_out"]) < 2e-4
    assert not net.are_sampling_parameters_initialized


def test_resample_smoke_shapes_like_reference_tests():
    """tests/samplers/test_resample.py:153-185 of the reference (identity VAE, zero eps) on the CUDA path."""
    from samplers_b200.inverse_problem import InverseProblem
    from samplers_b200.networks.base import EpsilonNetwork, LatentEpsilonNetwork
    from samplers_b200.noise import GaussianNoise
    from samplers_b200.operators import IdentityOperator
    from samplers_b200.samplers import ReSampleSampler

    class Mock(LatentEpsilonNetwork):
        def __init__(self, n=5):
            super().__init__(alphas_cumprod=torch.cat([torch.tensor([1.0]), torch.linspace(0.99, 0.5, n)]))

        def forward(self, x, t):
            return torch.zeros_like(x)

        @classmethod
        def from_pretrained(cls, *a, **k):
            raise NotImplementedError

        def get_latent_shape(self, x_shape):
            return x_shape

        def _encode(self, x, *, differentiable=False):
            return x

        def _decode(self, z, *, differentiable=False):
            return z * 1.0

        def set_sampling_parameters(self, num_sampling_steps, batch_size=1, num_reconstructions=1):
            self._batch_size = batch_size
            self.register_buffer("timesteps", torch.arange(1, num_sampling_steps + 1, device=self.alphas_cumprod.device))

        @property
        def is_condition_initialized(self):
            return True

    class Plain(EpsilonNetwork):
        def __init__(self):
            super().__init__(alphas_cumprod=torch.tensor([1.0, 0.9]))

        def forward(self, x, t):
            return torch.zeros_like(x)

        @classmethod
        def from_pretrained(cls, *a, **k):
            raise NotImplementedError

        def set_sampling_parameters(self, *a, **k):
            pass

        @property
        def is_condition_initialized(self):
            return True

    fast = dict(num_sampling_steps=5, max_optimization_iters=5, time_travel_interval=100)
    x_shape = (2, 4)
    for batch, R, expect in (((), 1, (1, 2, 4)), ((2,), 2, (2, 2, 2, 4))):
        prob = InverseProblem(operator=IdentityOperator(x_shape), observation=torch.randn(*batch, *x_shape, device=DEV),
                              noise=GaussianNoise(sigma=0.01))
        out = ReSampleSampler(Mock().to(DEV))(prob, num_reconstructions=R, **fast)
        assert out.shape == expect
    out = ReSampleSampler(Mock().to(DEV))(prob, decode_output=False, **fast)
    assert out.shape == (2, 1, 2, 4)
    with pytest.raises(TypeError, match="latent diffusion model"):
        ReSampleSampler(Plain())


def test_ddim_eps_step_and_stochastic_resample_kernels():
    from samplers_b200 import _native
    from samplers_b200.samplers.resample import ddim_eps_scalars, resample_scalars
    gen = torch.Generator().manual_seed(3)
    x, e, z = (torch.randn(2, 4, 9, 9, generator=gen) for _ in range(3))
    acp = torch.cat([torch.ones(1), torch.linspace(0.99, 0.2, 30)])
    for eta in (1.0, 0.0):
        ref_prev, ref_pred, ref_pseudo = ors.ddim_eps_step(x, e, acp, 20, 17, eta, z)
        sc = ddim_eps_scalars(acp, 20, 17, eta)
        prev, pred, pseudo = (torch.empty_like(x, device=DEV) for _ in range(3))
        _native.ddim_eps_step(x.to(DEV), e.to(DEV), z.to(DEV) if eta else None, sc, prev, pred, pseudo)
        assert rel_err(prev.cpu(), ref_prev) < 1e-6 and rel_err(pred.cpu(), ref_pred) < 1e-6
        assert rel_err(pseudo.cpu(), ref_pseudo) < 1e-6
    sigma = ors.compute_sigma(40.0, acp[20], acp[17])
    ref = ors.stochastic_resample(x, e, acp[17], sigma, z)
    out = torch.empty_like(x, device=DEV)
    _native.stochastic_resample(x.to(DEV), e.to(DEV), z.to(DEV), *resample_scalars(acp, 20, 17, 40.0), out)
    assert rel_err(out.cpu(), ref) < 1e-6


def test_adamw_kernel_matches_torch_optimizer():
    from samplers_b200 import _native
    gen = torch.Generator().manual_seed(4)
    p0 = torch.randn(1000, generator=gen)
    ref = p0.clone().requires_grad_()
    opt = torch.optim.AdamW([ref], lr=1e-2)
    p, m, v = p0.to(DEV), torch.zeros(1000, device=DEV), torch.zeros(1000, device=DEV)
    for step in range(1, 8):
        g = torch.randn(1000, generator=gen) * 0.1
        opt.zero_grad()
        ref.grad = g.clone()
        opt.step()
        _native.adamw_step(p, g.to(DEV), m, v, 1e-2, step)
        assert rel_err(p.cpu(), ref.detach()) < 1e-6


def test_pixel_optimization_on_device_loop_matches_oracle():
    """The flag-driven on-device AdamW loop stops on the same iteration as the reference's per-iteration check."""
    from samplers_b200 import operators as pops
    from samplers_b200.samplers.resample import ReSampleSampler
    from oracle.operators import OracleGaussianBlur
    shape = (3, 16, 16)
    gen = torch.Generator().manual_seed(5)
    x_true = torch.rand(2, *shape, generator=gen)
    ora = OracleGaussianBlur(shape, 9, 1.5)
    y = ora.apply(x_true)
    x_init = x_true + 0.02 * torch.randn(2, *shape, generator=gen)
    for eps, iters in ((0.015, 200), (1e-4, 40)):
        ref, ref_iters = ors.pixel_optimization(y, x_init, ora, eps, iters)
        op = pops.GaussianBlurOperator(shape, 9, 1.5).to(DEV)
        nat = op._native_cached(torch.device(DEV))
        ws = torch.empty(nat.workspace_bytes(2) // 4, device=DEV)
        got = ReSampleSampler._pixel_optimization(None, nat, y.reshape(2, -1).to(DEV).contiguous(), 1, ws,
                                                  x_init.reshape(2, -1).to(DEV).contiguous(), eps, iters)
        assert rel_err(got.cpu(), ref.reshape(2, -1)) < 1e-5, (eps, ref_iters)
