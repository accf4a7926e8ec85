"""bf16 state through the C ABI: K_bf16(inputs) == bf16_rn(K_fp32(float(inputs))) bit for bit, for K1
(identity, mask) and K2 (tensor noise, in-kernel Philox noise, fixed scale), vector and ragged sizes."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
BF = torch.bfloat16


@pytest.mark.parametrize("kind", ["identity", "mask"])
@pytest.mark.parametrize("shape", [(3, 32, 32), (3, 7, 9)])          # n % 8 == 0 and ragged
def test_k1_bf16_is_the_rounded_fp32_kernel(kind, shape):
    from samplers_b200 import _native, operators as P
    op = (P.IdentityOperator(shape) if kind == "identity" else P.RandomInpaintingOperator(shape, 0.7, seed=0, flatten=False)).to(DEV)
    nat = op._native_cached(torch.device(DEV))
    L, n = 5, nat.n
    gen = torch.Generator(device=DEV).manual_seed(0)
    x, e = (torch.randn(L, n, device=DEV, generator=gen).to(BF) for _ in range(2))
    y = torch.randn(2, n, device=DEV, generator=gen)
    if kind == "mask":
        y = y * (~op.mask).float().reshape(1, -1).to(DEV)
    sa, s1, w = 0.8366600275039673, 0.547722578048706, 400.0
    cot32, part32 = torch.empty(L, n, device=DEV), torch.empty(L, nat.err_parts, device=DEV)
    _native.dps_pre(nat, x.float(), e.float(), y, 3, sa, s1, w, cot32, part32, None)
    cot16, part16 = torch.empty(L, n, device=DEV, dtype=BF), torch.empty(L, nat.err_parts, device=DEV)
    _native.dps_pre_bf16(nat, x, e, y, 3, sa, s1, w, cot16, part16)
    assert torch.equal(cot16, cot32.to(BF))
    # the partial sums are fp32 sums of the same fp32 residuals (the split into parts may differ)
    assert torch.allclose(part16.sum(1), part32.sum(1), rtol=1e-6)
    row = torch.tensor([[sa, s1, float(torch.tensor(w) / torch.tensor(sa)), 0, 0, 0, 0, 0]], device=DEV)
    cot_dev = torch.empty_like(cot16)
    _native.dps_pre_bf16(nat, x, e, y, 3, 1.0, 0.0, 1.0, cot_dev, part16, step_row=row)
    assert torch.equal(cot_dev, cot16)


@pytest.mark.parametrize("n", [3 * 32 * 32, 1001])
@pytest.mark.parametrize("noise", ["tensor", "philox", "none", "fixed"])
def test_k2_bf16_is_the_rounded_fp32_kernel(n, noise):
    from samplers_b200 import _native
    L, parts = 3, 64
    gen = torch.Generator(device=DEV).manual_seed(1)
    x, e, c, v, z = (torch.randn(L, n, device=DEV, generator=gen).to(BF) for _ in range(5))
    part = torch.rand(L, parts, device=DEV, generator=gen)
    sc = dict(sa=0.83666, s1=0.54772, c_ell=0.97, c_s=0.021, std=0.0 if noise == "none" else 0.11, gamma=1.3)
    f = [t.float() for t in (x, e, c, v)]
    out32, err32 = torch.empty(L, n, device=DEV), torch.empty(L, device=DEV)
    out16, err16 = torch.empty(L, n, device=DEV, dtype=BF), torch.empty(L, device=DEV)
    fixed = noise == "fixed"
    ep, np_ = (None, 0) if fixed else (part, parts)
    if noise == "philox":
        _native.dps_post_philox(*f, ep, np_, n, *sc.values(), 77, 5, out32, err32)
        _native.dps_post_bf16(x, e, c, v, None, ep, np_, n, *sc.values(), out16, err16, philox=(77, 5))
        ss = torch.tensor([77, 5], dtype=torch.int64, device=DEV)
        row = torch.tensor([[sc["sa"], sc["s1"], 0.0, sc["c_ell"], sc["c_s"], sc["std"], sc["gamma"], 0.0]], device=DEV)
        out_dev = torch.empty_like(out16)
        _native.dps_post_bf16(x, e, c, v, None, ep, np_, n, 1, 0, 0, 0, 0, 0, out_dev, None, step_row=row, seed_step=ss)
        assert torch.equal(out_dev, out16)
    else:
        zz = None if noise == "none" else z
        _native.dps_post(*f, None if zz is None else zz.float(), ep, np_, n, *sc.values(), out32, None if fixed else err32)
        _native.dps_post_bf16(x, e, c, v, zz, ep, np_, n, *sc.values(), out16, None if fixed else err16)
    assert torch.equal(out16, out32.to(BF))
    if not fixed:
        assert torch.equal(err16, err32)
    # in place, as the graphed step uses it
    xa = x.clone()
    if noise != "philox":
        _native.dps_post_bf16(xa, e, c, v, None if noise == "none" else z, ep, np_, n, *sc.values(), xa, None)
        assert torch.equal(xa, out16)


@pytest.mark.parametrize("shape", [(3, 32, 32), (1, 8, 20)])
def test_k1_box_bf16_is_the_rounded_fp32_kernel(shape):
    from samplers_b200 import _native, operators as P
    op = P.BoxDownsampleOperator(shape, 4).to(DEV)
    nat = op._native_cached(torch.device(DEV))
    L, n = 4, nat.n
    gen = torch.Generator(device=DEV).manual_seed(2)
    x, e = (torch.randn(L, n, device=DEV, generator=gen).to(BF) for _ in range(2))
    y = torch.randn(2, nat.n_y, device=DEV, generator=gen)
    sa, s1, w = 0.8366600275039673, 0.547722578048706, 400.0
    cot32, part32 = torch.empty(L, n, device=DEV), torch.empty(L, nat.err_parts, device=DEV)
    _native.dps_pre(nat, x.float(), e.float(), y, 2, sa, s1, w, cot32, part32, None)
    cot16, part16 = torch.empty(L, n, device=DEV, dtype=BF), torch.empty(L, nat.err_parts, device=DEV)
    _native.dps_pre_bf16(nat, x, e, y, 2, sa, s1, w, cot16, part16)
    assert torch.equal(cot16, cot32.to(BF))
    assert torch.allclose(part16.sum(1), part32.sum(1), rtol=1e-6)


@pytest.mark.parametrize("shape,ksize,split", [((3, 256, 256), 61, None), ((3, 256, 256), 61, "2"), ((1, 512, 512), 61, None),
                                               ((2, 64, 96), 9, None)])
def test_k1_blur_bf16_is_the_rounded_fp32_kernel(psx_env, shape, ksize, split):
    """Separable blur on a bf16 state: bf16 x_t / eps in, fp32 intermediates, bf16 cotangent out.  The identity is
    between the bf16 strip kernels and the fp32 strip kernels (PSX_NO_TC: the fp32 default for 256 x 256 planes is the
    tensor-core kernel, a different summation order)."""
    from samplers_b200 import _native, operators as P
    psx_env(PSX_SPLIT=split, PSX_NO_TC="1")
    op = P.GaussianBlurOperator(shape, ksize, 3.0 if ksize == 61 else 1.5).to(DEV)
    nat = op._native_cached(torch.device(DEV))
    L, n = 4, nat.n
    gen = torch.Generator(device=DEV).manual_seed(3)
    x, e = (torch.randn(L, n, device=DEV, generator=gen).to(BF) for _ in range(2))
    y = torch.randn(2, nat.n_y, device=DEV, generator=gen)
    sa, s1, w = 0.8366600275039673, 0.547722578048706, 400.0
    ws = torch.empty(nat.workspace_bytes(L) // 4, device=DEV)
    cot32, part32 = torch.empty(L, n, device=DEV), torch.empty(L, nat.err_parts, device=DEV)
    _native.dps_pre(nat, x.float(), e.float(), y, 2, sa, s1, w, cot32, part32, ws)
    cot16, part16 = torch.empty(L, n, device=DEV, dtype=BF), torch.empty(L, nat.err_parts, device=DEV)
    _native.dps_pre_bf16(nat, x, e, y, 2, sa, s1, w, cot16, part16, ws=ws)
    assert torch.equal(cot16, cot32.to(BF))
    assert torch.equal(part16, part32)


def test_bf16_k1_rejects_operators_without_a_bf16_kernel():
    from samplers_b200 import _native, operators as P
    op = P.MotionBlurOperator((3, 32, 32), kernel_size=9, angle_deg=30.0).to(DEV)
    nat = op._native_cached(torch.device(DEV))
    x = torch.zeros(2, nat.n, device=DEV, dtype=BF)
    with pytest.raises(NotImplementedError):
        _native.dps_pre_bf16(nat, x, x, torch.zeros(1, nat.n_y, device=DEV), 2, 0.8, 0.6, 1.0, torch.empty_like(x),
                             torch.empty(2, nat.err_parts, device=DEV))


@pytest.mark.parametrize("graph", [False, True])
@pytest.mark.parametrize("net_dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("op_kind", ["mask", "blur"])
def test_sampler_on_a_bf16_state(graph, net_dtype, op_kind):
    """DPSSampler(state_dtype=bfloat16): the run stays close to the fp32-state run of the same network (stated
    tolerance: bf16 rounding of the stored state, a few 2^-9 per step), eager and graph, fp32 and bf16 networks."""
    from samplers_b200 import operators as P
    from samplers_b200.inverse_problem import InverseProblem
    from samplers_b200.networks import DDPMNetwork
    from samplers_b200.noise import GaussianNoise
    from samplers_b200.samplers import DPSSampler
    from tests._golden import rel_err
    shape = (3, 32, 32)
    op = (P.RandomInpaintingOperator(shape, 0.5, seed=0, flatten=False) if op_kind == "mask"
          else P.GaussianBlurOperator(shape, 9, 1.5)).to(DEV)
    gen = torch.Generator(device=DEV).manual_seed(1)
    x_true = torch.rand(shape, device=DEV, generator=gen) * 2 - 1
    keep = (~op.mask).float().to(DEV) if op_kind == "mask" else 1.0
    y = op.apply(x_true[None])[0] + 0.05 * torch.randn(shape, device=DEV, generator=gen) * keep
    prob = InverseProblem(operator=op, observation=y, noise=GaussianNoise(sigma=0.05))
    torch.manual_seed(1234)
    net = DDPMNetwork.from_config("tiny", device=DEV, torch_dtype=net_dtype)
    x_init = torch.randn(2, *shape, device=DEV, generator=torch.Generator(device=DEV).manual_seed(5))
    outs = {}
    for sd in (torch.float32, torch.bfloat16):
        s = DPSSampler(net, cuda_graph=graph, philox_seed=9, state_dtype=sd)
        s.draw = lambda shape_, device, dtype: x_init.clone()
        run = s.prepare(prob, num_sampling_steps=8, num_reconstructions=2, gamma=0.05)
        try:
            if graph:
                run.capture()
            for k in range(3):
                run.step(k)
            assert run.x.dtype == sd and run.cot.dtype == sd
            outs[sd] = run.x.float().clone()
            fin = run.finalize()
            assert fin.dtype == torch.float32 and torch.isfinite(fin).all()
        finally:
            s.release()
    assert rel_err(outs[torch.bfloat16].cpu(), outs[torch.float32].cpu()) < 3e-2


def test_bf16_state_needs_a_pointwise_operator():
    from samplers_b200 import operators as P
    from samplers_b200.inverse_problem import InverseProblem
    from samplers_b200.networks import DDPMNetwork
    from samplers_b200.noise import GaussianNoise
    from samplers_b200.samplers import DPSSampler
    shape = (3, 32, 32)
    op = P.MotionBlurOperator(shape, kernel_size=9, angle_deg=30.0).to(DEV)
    prob = InverseProblem(operator=op, observation=torch.zeros(shape, device=DEV), noise=GaussianNoise(sigma=0.05))
    net = DDPMNetwork.from_config("tiny", device=DEV)
    with pytest.raises(NotImplementedError):
        DPSSampler(net, state_dtype=BF)(prob, num_sampling_steps=5)
    assert not net.are_sampling_parameters_initialized
