"""Two-GPU checks (skipped on a single-GPU box): PSLD / ReSample sharded over two ranks with all-reduced
batch-global norms (``process_group``) must reproduce the reference's FULL-batch recording, sample for sample."""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, kind, name, q):
    import torch.distributed as dist
    from tests._golden import (PsldGolden, ResampleGolden, make_latent_network, make_psld_problem,
                               make_resample_problem, rel_err)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        from samplers_b200.samplers import PSLDSampler, ReSampleSampler
        if kind == "psld":
            g = PsldGolden(name)
            m = g.meta
            net, prob = make_latent_network(g, dev), make_psld_problem(g, dev)
            draws = iter([g["z_init"]] + [g["noise"][k] for k in range(g.K)])
            s = PSLDSampler(net, process_group=dist.group.WORLD)
            kw = dict(num_sampling_steps=m["steps"], gamma=m["gamma"], omega=m["omega"], eta=m["eta"])
        else:
            g = ResampleGolden(name)
            m = g.meta
            net, prob = make_latent_network(g, dev), make_resample_problem(g, dev)
            draws = iter(g["draws"])
            s = ReSampleSampler(net, process_group=dist.group.WORLD)
            kw = dict(num_sampling_steps=m["steps"], **m["kw"])
        per = m["R"] // world
        # this rank's rows of every recorded full-batch draw
        s.draw = lambda shape, device, dtype: next(draws)[rank * per:(rank + 1) * per].to(device)
        out = s(prob, num_reconstructions=per, **kw).cpu()
        want = g["x_out"][rank * per:(rank + 1) * per]
        q.put((rank, tuple(out.shape), tuple(want.shape), rel_err(out, want)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("kind,name", [("psld", "identity"), ("psld", "box2"), ("resample", "identity")])
def test_sharded_run_with_global_norms_reproduces_full_batch_reference(kind, name):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, kind, name, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    for rank, shape, want_shape, err in results:
        assert shape == want_shape
        assert err < 2e-4, (rank, err)
