"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: sharding + the terminal exchange."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from samplers_b200.distributed import combine_posterior, shard_count


def test_shard_count_partitions_exactly():
    for total in (1, 7, 8, 16, 255, 256):
        for world in (1, 2, 3, 8):
            spans = [shard_count(total, r, world) for r in range(world)]
            assert sum(c for _, c in spans) == total
            assert all(spans[i][0] + spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            assert max(c for _, c in spans) - min(c for _, c in spans) <= 1
    with pytest.raises(ValueError):
        shard_count(4, 2, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, counts, n, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(100 + rank)
        local = torch.randn(counts[rank], n, generator=g)
        mx = max(counts)
        gathered = torch.zeros(world * mx, n)
        slot = gathered[rank * mx: rank * mx + counts[rank]]
        slot.copy_(local)
        out = combine_posterior(slot, gathered, local.sum(0), local.square().sum(0), counts)
        q.put((rank, out.samples, out.mean, out.variance))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("counts", [[3, 3], [4, 3]])
def test_terminal_gather_and_moments_world2(counts):
    world, n = 2, 48
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, counts, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = sorted((q.get(timeout=120) for _ in range(world)), key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    expected = torch.cat([torch.randn(counts[r], n, generator=torch.Generator().manual_seed(100 + r))
                          for r in range(world)])
    for _, samples, mean, var in results:  # every rank ends with the same, complete answer
        assert torch.equal(samples, expected)
        assert torch.allclose(mean, expected.mean(0), atol=1e-6)
        assert torch.allclose(var, expected.var(0, unbiased=True), atol=1e-5)


def test_combine_posterior_single_process():
    local = torch.randn(5, 12, generator=torch.Generator().manual_seed(0))
    gathered = local.clone()
    out = combine_posterior(gathered, gathered, local.sum(0), local.square().sum(0), [5])
    assert torch.equal(out.samples, local)
    assert torch.allclose(out.variance, local.var(0), atol=1e-6)
