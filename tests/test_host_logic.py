"""CPU tests of the host side: C-ABI surface, planner, shapes, error conventions, no-fallback rules."""
import os
import re
import subprocess
import sys

import pytest
import torch

from oracle import dps as odps
from tests._golden import Golden, golden_names

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    text = open(os.path.join(ROOT, "include", "psx.h")).read()
    return sorted(set(re.findall(r"PSX_API\s+[\w\s\*]+?\b(psx_\w+)\s*\(", text)))


def test_library_exports_every_header_symbol():
    from samplers_b200 import _native
    from samplers_b200.build import build_native
    build_native()
    lib = _native.load()
    syms = _header_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), s
    assert sorted(_native.PROTOTYPES) == syms  # the ctypes table binds exactly the header
    out = subprocess.run(["nm", "-D", "--defined-only", _native.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r"\bT (psx_\w+)", out))
    assert exported == set(syms)
    assert lib.psx_abi_version() == _native.ABI_VERSION


def test_abi_argument_validation_without_gpu():
    """Descriptor creation / validation is host-side: errors come back as codes + messages, never crashes."""
    import ctypes as C
    from samplers_b200 import _native
    lib = _native.load()
    h = C.c_void_p()
    assert lib.psx_op_create_identity(0, C.byref(h)) == _native.PSX_ERR_INVALID
    assert b"positive" in lib.psx_last_error()
    assert lib.psx_op_create_box(3, 10, 10, 4, C.byref(h)) == _native.PSX_ERR_INVALID
    taps = (C.c_float * 4)(0.25, 0.25, 0.25, 0.25)
    assert lib.psx_op_create_sepblur(3, 8, 8, taps, 4, taps, 4, C.byref(h)) == _native.PSX_ERR_INVALID  # even taps
    assert lib.psx_op_create_identity(48, C.byref(h)) == _native.PSX_OK
    assert lib.psx_op_x_numel(h) == 48 and lib.psx_op_kind(h) == _native.OP_IDENTITY
    assert lib.psx_op_workspace_bytes(h, 4) == 0 and lib.psx_op_err_parts(h) >= 1
    assert lib.psx_dps_pre(h, None, None, None, 1, 1, 1.0, 0.0, 1.0, None, None, None, None, 0, None) == \
        _native.PSX_ERR_INVALID
    # the bridge-mean pair (ABI 5): offered by the tensor-core blur only; argument checks come before any CUDA call
    assert lib.psx_op_fuses_mean(h, 16) == 0 and lib.psx_op_fuses_mean(None, 16) == 0
    assert lib.psx_dps_pre_mean(h, None, None, None, 1, 1, 1.0, 0.0, 1.0, 0.5, 0.5, None, 0.0, None, None, None, None,
                                0, None) == _native.PSX_ERR_INVALID
    assert lib.psx_dps_post_mean(None, None, None, None, None, 0, 1, 48, 0.6, 0.0, 1.0, None, None, None) == \
        _native.PSX_ERR_INVALID
    assert lib.psx_dps_post_mean_dev(None, None, None, None, None, 0, 1, 48, None, None, None, None) == \
        _native.PSX_ERR_INVALID
    assert lib.psx_op_destroy(h) == _native.PSX_OK
    taps3 = (C.c_float * 3)(0.25, 0.5, 0.25)
    assert lib.psx_op_create_sepblur(3, 8, 8, taps3, 3, taps3, 3, C.byref(h)) == _native.PSX_OK
    assert lib.psx_op_workspace_bytes(h, 2) == 2 * (2 * 3 * 8 * 8 * 4)   # h1 and h2 of the CUDA-core K1: two regions
    lib.psx_op_destroy(h)


@pytest.mark.parametrize("name", golden_names())
def test_step_planner_matches_oracle_and_reference_schedule(name):
    from samplers_b200.samplers.utils.bridge_kernels import plan_steps
    g = Golden(name)
    plan = plan_steps(g["acp"], g["timesteps"].tolist(), g.meta["eta"])
    assert [p.t for p in plan] == g.meta["t"] and [p.t_prev for p in plan] == g.meta["t_prev"]
    for p in plan:
        c_ell, c_s, std = odps.bridge_coefficients(g["acp"], p.t, p.t_prev, g.meta["s"], g.meta["eta"])
        assert p.c_ell == float(c_ell.float()) and p.c_s == float(c_s.float()) and p.std == float(std.float())
        a = g["acp"][p.t]
        assert p.sqrt_acp == float(a ** 0.5) and p.sqrt_1m_acp == float((1 - a) ** 0.5)


def test_batch_view_shapes():
    """tests/samplers/test_batch_view.py:15-22 of the reference + the rest of the surface."""
    from samplers_b200.samplers.utils import BatchView
    v = BatchView((2, 3), 4, (3, 8, 8))
    assert v.shape == (2, 3, 4, 3, 8, 8) and v.flat_shape == (24, 3, 8, 8)
    assert v.leading_shape == (2, 3, 4) and v.leading_size == 24 and v.batch_size == 6
    assert v.per_sample_broadcast_shape == (24, 1, 1, 1)
    assert BatchView((), 5, (7,)).shape == (5, 7) and BatchView(2, 1, 3).shape == (2, 1, 3)
    y = torch.arange(6 * 5.0).reshape(2, 3, 5)
    rep = BatchView((2, 3), 4, (5,)).repeat_observation(y)
    assert rep.shape == (24, 5)
    assert torch.equal(rep[0], y[0, 0]) and torch.equal(rep[3], y[0, 0]) and torch.equal(rep[4], y[0, 1])
    x = torch.randn(24, 5)
    assert torch.equal(BatchView((2, 3), 4, (5,)).flatten(BatchView((2, 3), 4, (5,)).unflatten(x)), x)


def test_operator_shapes_masks_and_errors():
    from samplers_b200 import operators as pops
    shape = (3, 16, 16)
    assert pops.IdentityOperator(shape).y_shape == shape
    assert pops.IdentityOperator(shape, flatten=True).y_shape == (768,)
    x = torch.randn(2, *shape)
    ident = pops.IdentityOperator(shape, flatten=True)
    assert torch.equal(ident.apply_transpose(ident.apply(x)), x)  # pure views: no kernel involved
    c = pops.CenterInpaintingOperator(shape, 0.5)
    assert c.shape == (3 * (256 - 64), 768) and c.y_shape == (576,) and c.mask[:, 4:12, 4:12].all()
    o = pops.CenterOutpaintingOperator(shape, 0.5)
    assert o.shape[0] == 3 * 64
    s = pops.SidePaintingOperator(shape, 0.25, left=True)
    assert s.mask[..., :4].all() and not s.mask[..., 4:].any()
    r = pops.RandomInpaintingOperator(shape, 0.7, seed=1)
    assert 0.6 < r.mask.float().mean() < 0.8
    assert pops.BoxDownsampleOperator(shape, 4).y_shape == (3, 4, 4)
    assert pops.GaussianBlurOperator(shape).y_shape == shape
    assert abs(float(pops.gaussian_taps(61, 3.0).double().sum()) - 1) < 1e-6
    for bad in (lambda: pops.CenterInpaintingOperator(shape, 1.5), lambda: pops.SidePaintingOperator(shape, -0.1),
                lambda: pops.InpaintingOperator(shape, torch.zeros(3, 16, 15, dtype=torch.bool)),
                lambda: pops.BoxDownsampleOperator(shape, 5), lambda: pops.GaussianBlurOperator(shape, 60),
                lambda: pops.SeparableBlurOperator(shape, torch.ones(4)),
                lambda: pops.get_mask_inpaint_center(shape, 0.8, 0.2)):
        with pytest.raises(ValueError):
            bad()

    class NoAdjoint(pops.LinearOperator):
        def _infer_y_shape(self, x_shape, device=None):
            return x_shape

        def apply(self, x):
            return x

    with pytest.raises(NotImplementedError):
        NoAdjoint(shape).apply_transpose(x)
    with pytest.raises(NotImplementedError):
        NoAdjoint(shape).apply_pseudo_inverse(x)


def test_noise_models_host_behaviour():
    from samplers_b200.noise import GaussianNoise, PoissonNoise
    gn, pn = GaussianNoise(sigma=0.05), PoissonNoise(rate=4.0)
    assert gn.device.type == "cpu" and gn.dtype == torch.float32
    assert abs(gn._likelihood_weight() - odps.likelihood_weight("gaussian", torch.tensor(0.05))) < 1e-9
    assert abs(pn._likelihood_weight() - odps.likelihood_weight("poisson", torch.tensor(4.0))) < 1e-12
    r = torch.randn(3, 7)
    assert torch.allclose(gn.log_prob(r), odps.log_prob(r, "gaussian", torch.tensor(0.05)))
    assert torch.allclose(pn.log_prob(r), odps.log_prob(r, "poisson", torch.tensor(4.0)))
    rr = r.clone().requires_grad_()
    (ag,) = torch.autograd.grad(gn.log_prob(rr).sum(), rr)
    assert torch.allclose(gn.score(r), ag, rtol=1e-5)
    g = torch.Generator().manual_seed(0)
    assert gn.sample((4, 5), generator=g).shape == (4, 5)
    k = pn.sample((1000,), generator=g) + 4.0
    assert torch.all(k >= 0) and torch.allclose(k, k.round())
    for bad in (lambda: GaussianNoise(sigma=-1.0), lambda: PoissonNoise(rate=0.0),
                lambda: GaussianNoise(sigma=torch.ones(3))):
        with pytest.raises(ValueError):
            bad()


def test_no_cpu_fallback_and_no_oracle_in_product():
    from samplers_b200 import operators as pops
    from samplers_b200.inverse_problem import InverseProblem
    from samplers_b200.networks import DDPMNetwork
    from samplers_b200.noise import GaussianNoise
    from samplers_b200.samplers import DPSSampler
    shape = (3, 32, 32)
    op = pops.GaussianBlurOperator(shape, 9, 1.5)
    with pytest.raises(RuntimeError, match="CUDA"):
        op.apply(torch.zeros(1, *shape))
    net = DDPMNetwork.from_config("tiny")
    prob = InverseProblem(operator=pops.IdentityOperator(shape), observation=torch.zeros(shape),
                          noise=GaussianNoise(sigma=0.1))
    with pytest.raises(RuntimeError, match="CUDA"):
        DPSSampler(net)(prob, num_sampling_steps=4)
    assert not net.are_sampling_parameters_initialized  # cleaned up even though it raised
    with pytest.raises(RuntimeError, match="set_sampling_parameters"):
        net.forward(torch.zeros(1, *shape), 0)
    for dirpath, _, files in os.walk(os.path.join(ROOT, "samplers_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), f"{f} imports the oracle"


def test_network_contract():
    from samplers_b200.networks import DDPMNetwork
    from oracle.schedule import ddpm_linear_alphas_cumprod, leading_timesteps_ascending, padded_clipped_acp
    net = DDPMNetwork.from_config("tiny")
    assert torch.equal(net.alphas_cumprod, padded_clipped_acp(ddpm_linear_alphas_cumprod()))
    assert float(net.alphas_cumprod[0]) == 1.0 and net.alphas_cumprod.numel() == 1001
    net.set_sampling_parameters(50)
    assert torch.equal(net.timesteps, leading_timesteps_ascending(50)) and net.are_sampling_parameters_initialized
    net.clear_sampling_parameters()
    assert not net.are_sampling_parameters_initialized
    with pytest.raises(ImportError):
        DDPMNetwork.from_pretrained("google/ddpm-celebahq-256")
    full = sum(p.numel() for p in DDPMNetwork.from_config("google/ddpm-celebahq-256").unet.parameters())
    assert full == 113_673_219  # parameter count of the published ddpm-celebahq-256 UNet2DModel


def test_bench_reference_arm_runs_nothing_on_nonzero_rank():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference"],
                         capture_output=True, text=True, env=env, timeout=120)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_step_rows_hold_the_by_value_scalars():
    """Device step table (psx_dps_*_dev): fp32 copies of the plan, w / sa as ONE fp32 division (as libpsx does)."""
    import numpy as np
    from oracle.schedule import ddpm_linear_alphas_cumprod, leading_timesteps_ascending, padded_clipped_acp
    from samplers_b200.samplers.utils.bridge_kernels import STEP_ROW, plan_steps, step_rows
    acp = padded_clipped_acp(ddpm_linear_alphas_cumprod())
    plan = plan_steps(acp, leading_timesteps_ascending(20).tolist(), eta=1.0)
    rows = step_rows(plan, 400.0, 1.5)
    assert rows.shape == (len(plan), STEP_ROW) and rows.dtype == torch.float32
    for k, sc in enumerate(plan):
        want = [sc.sqrt_acp, sc.sqrt_1m_acp, float(np.float32(400.0) / np.float32(sc.sqrt_acp)), sc.c_ell, sc.c_s,
                sc.std, 1.5, 0.0]
        assert rows[k].tolist() == [float(np.float32(v)) for v in want]
    per_step = step_rows(plan, 2.0, lambda sc: 0.5 * sc.sqrt_1m_acp)
    assert per_step[3, 6].item() == float(np.float32(0.5 * plan[3].sqrt_1m_acp))


def test_stable_diffusion_network_surface():
    """StableDiffusionNetwork mirrors samplers/networks/diffusers/stable_diffusion.py: latent shape rule, ascending
    timesteps, conditioning life cycle, CFG batch doubling, VAE scaling factor -- on the narrow random-init config."""
    from samplers_b200.networks import LatentEpsilonNetwork, StableDiffusionCondition, StableDiffusionNetwork
    from samplers_b200.networks.sd15 import SD15, AutoencoderKLLite, UNet2DConditionLite
    with torch.device("meta"):    # the full SD-1.5 shapes, without allocating them
        unet, vae = UNet2DConditionLite(**SD15["unet"]), AutoencoderKLLite(**SD15["vae"])
    assert sum(p.numel() for p in unet.parameters()) == 859_520_964      # the published SD-1.5 UNet size
    assert abs(sum(p.numel() for p in vae.parameters()) - 83_653_863) < 4096
    net = StableDiffusionNetwork.from_config("sd15-tiny")
    assert isinstance(net, LatentEpsilonNetwork) and net.dtype == torch.float32
    assert float(net.alphas_cumprod[0]) == 1.0 and net.alphas_cumprod.numel() == 1001
    with pytest.raises(RuntimeError):
        net.set_condition(None)                      # before set_sampling_parameters, as the reference
    with pytest.raises(ValueError):
        net.get_latent_shape((3, 60, 64))
    assert net.get_latent_shape((3, 64, 64)) == (4, 8, 8)
    net.set_sampling_parameters(10, batch_size=1, num_reconstructions=2)
    assert net.timesteps.tolist() == sorted(net.timesteps.tolist()) and net.are_sampling_parameters_initialized
    z = torch.randn(2, 4, 8, 8, generator=torch.Generator().manual_seed(0))
    with pytest.raises(RuntimeError):
        net.forward(z, 500)                          # before set_condition
    net.set_condition(None)
    e0 = net.forward(z, 500)
    assert e0.shape == z.shape and torch.equal(e0, net.forward(z, torch.tensor([500])))
    x = net.decode(z)
    assert x.shape == (2, 3, 64, 64) and not x.requires_grad
    assert net.encode(x).shape == z.shape
    zg = z.clone().requires_grad_()
    assert net.decode(zg, differentiable=True).requires_grad
    # decode(z) = vae.decode(z / scaling_factor); encode = mean * scaling_factor
    assert torch.allclose(net.vae.decode(z / 0.18215), x, atol=1e-6)
    emb = torch.randn(1, 7, 32, generator=torch.Generator().manual_seed(1))
    net.set_condition(StableDiffusionCondition(guidance_scale=3.0, prompt_embeds=emb))
    assert net._conditioning.prompt_embeds.shape == (4, 7, 32) and net._conditioning.do_classifier_free_guidance
    e1 = net.forward(z, 500)
    net.set_condition(StableDiffusionCondition(guidance_scale=1.0, prompt_embeds=emb))
    e_text = net.forward(z, 500)
    assert torch.allclose(e1, e0 + 3.0 * (e_text - e0), atol=1e-5)       # CFG formula
    with pytest.raises(NotImplementedError):
        net.set_condition(StableDiffusionCondition(prompt="a photo"))      # no text encoder in this build
    with pytest.raises(ValueError):
        net.set_condition(StableDiffusionCondition(prompt_embeds=torch.zeros(1, 5, 32)))
    net.clear_condition()
    net.clear_sampling_parameters()
    assert not net.is_condition_initialized and not net.are_sampling_parameters_initialized
