"""In-kernel noise through the C ABI: psx_philox_normal against oracle/philox.py (Random123-pinned),
psx_dps_post_philox bit-equal to psx_dps_post fed with that field, eager == graph replay."""
import numpy as np
import pytest
import torch

from oracle.philox import normals
from tests._golden import rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("numel,seed,step", [(4096, 0, 0), (3 * 64 * 64 * 5, 1234, 7), (1003, 2 ** 40 + 17, 2 ** 33 + 5)])
def test_philox_field_matches_oracle(numel, seed, step):
    from samplers_b200 import _native
    out = torch.empty(numel, device=DEV)
    _native.philox_normal(out, seed, step)
    want = normals(numel, seed, step)
    got = out.cpu().numpy()
    # same Philox bits; fp32 log / sincospi vs the oracle's fp64 transform
    assert np.max(np.abs(got - want) / np.maximum(1.0, np.abs(want))) < 5e-6


@pytest.mark.parametrize("n", [3 * 32 * 32, 1001])
@pytest.mark.parametrize("fixed", [False, True])
def test_post_philox_equals_post_with_the_same_field(n, fixed):
    from samplers_b200 import _native
    L, parts = 3, 64
    gen = torch.Generator(device=DEV).manual_seed(0)
    x, e, c, v = (torch.randn(L, n, device=DEV, generator=gen) for _ in range(4))
    part = torch.rand(L, parts, device=DEV, generator=gen)
    z = torch.empty(L, n, device=DEV)
    _native.philox_normal(z, 99, 3)
    sc = dict(sa=0.83666, s1=0.54772, c_ell=0.97, c_s=0.021, std=0.11, gamma=1.3)
    a, b = torch.empty(L, n, device=DEV), torch.empty(L, n, device=DEV)
    ea, eb = torch.empty(L, device=DEV), torch.empty(L, device=DEV)
    if fixed:
        _native.dps_post(x, e, c, v, z, None, 0, n, *sc.values(), a, None)
        _native.dps_post_philox(x, e, c, v, None, 0, n, *sc.values(), 99, 3, b, None)
    else:
        _native.dps_post(x, e, c, v, z, part, parts, n, *sc.values(), a, ea)
        _native.dps_post_philox(x, e, c, v, part, parts, n, *sc.values(), 99, 3, b, eb)
        assert torch.equal(ea, eb)
    assert torch.equal(a, b)
    # device-scalar form, {seed, step} from device memory
    row = torch.tensor([[sc["sa"], sc["s1"], 0.0, sc["c_ell"], sc["c_s"], sc["std"], sc["gamma"], 0.0]], device=DEV)
    ss = torch.tensor([99, 3], dtype=torch.int64, device=DEV)
    d = torch.empty(L, n, device=DEV)
    _native.dps_post_philox_dev(x, e, c, v, None if fixed else part, 0 if fixed else parts, n, row, ss, d, None)
    assert torch.equal(d, a)


def test_sampler_with_philox_noise_eager_equals_graph():
    from samplers_b200.samplers import DPSSampler
    from tests.test_gpu_graph import SHAPE, _problem, _tiny_net
    net, prob = _tiny_net(), _problem("blur")
    x_init = torch.randn(2, *SHAPE, device=DEV, generator=torch.Generator(device=DEV).manual_seed(5))
    outs = []
    for graph in (False, True):
        s = DPSSampler(net, cuda_graph=graph, philox_seed=4242)
        s.draw = lambda shape, device, dtype: x_init.clone()     # only the initial state; steps use Philox
        run = s.prepare(prob, num_sampling_steps=8, num_reconstructions=2)
        try:
            if graph:
                run.capture()
            run.step(0)
            first = run.x.clone()
            for k in range(1, run.num_steps):
                run.step(k)
            outs.append((first, run.x.clone()))
        finally:
            s.release()
    # same (seed, step) field in both modes: the first step agrees to backward-pass noise, the run stays close
    assert rel_err(outs[1][0].cpu(), outs[0][0].cpu()) < 2e-4
    assert rel_err(outs[1][1].cpu(), outs[0][1].cpu()) < 5e-3
    assert torch.isfinite(outs[1][1]).all()
    # and the noise really is the oracle's field: redo step 0 by hand with injected z
    s = DPSSampler(net)
    s.draw = lambda shape, device, dtype: x_init.clone()
    run = s.prepare(prob, num_sampling_steps=8, num_reconstructions=2)
    try:
        z = torch.from_numpy(normals(run.L * run.n, 4242, 0)).to(DEV).view(run.L, run.n)
        run.step(0, z=z)
        assert rel_err(run.x.cpu(), outs[0][0].cpu()) < 2e-4
    finally:
        s.release()
