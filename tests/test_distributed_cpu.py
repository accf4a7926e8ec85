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


def _run_world(target, world, make_args):
    """Spawn `world` gloo ranks of `target(rank, world, port, *make_args(q))` and return their queue results sorted by
    rank.  The rendezvous port is probed and released before the ranks bind it, so another process can take it in
    between: a failed rendezvous is retried on a fresh port (twice) instead of failing the suite."""
    import queue
    import time
    ctx = mp.get_context("spawn")
    for _ in range(3):
        q = ctx.Queue()
        port = _free_port()
        procs = [ctx.Process(target=target, args=(r, world, port, *make_args(q))) for r in range(world)]
        for p in procs:
            p.start()
        results, deadline = [], time.time() + 120
        while len(results) < world and time.time() < deadline:
            try:
                results.append(q.get(timeout=1))
            except queue.Empty:
                if any(p.exitcode not in (None, 0) for p in procs):
                    break
        for p in procs:
            p.join(timeout=60)
            if p.is_alive():
                p.terminate()
        if len(results) == world and all(p.exitcode == 0 for p in procs):
            return sorted(results, key=lambda t: t[0])
    raise AssertionError("gloo world did not complete in three attempts")


def _draw(rank, count, n, offset, scale):
    return offset + scale * torch.randn(count, n, generator=torch.Generator().manual_seed(100 + rank))


def _worker(rank, world, port, counts, n, q, offset=0.0, scale=1.0):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        local = _draw(rank, counts[rank], n, offset, scale)
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
    results = _run_world(_worker, world, lambda q: (counts, n, q))
    expected = torch.cat([_draw(r, counts[r], n, 0.0, 1.0) for r in range(world)])
    for _, samples, mean, var in results:  # every rank ends with the same, complete answer
        assert torch.equal(samples, expected)
        assert torch.allclose(mean, expected.mean(0), atol=1e-6)
        assert torch.allclose(var, expected.var(0, unbiased=True), atol=1e-5)


def test_posterior_variance_of_a_converged_posterior_world2():
    """256 samples with mean 0.9 and standard deviation 1e-3: sum(x^2) - R mean^2 has no correct digit left in fp32
    (0.81 * 256 against a spread of 2.6e-4); the two-pass form keeps 4+ digits."""
    world, n, counts = 2, 32, [128, 128]
    results = _run_world(_worker, world, lambda q: (counts, n, q, 0.9, 1e-3))
    expected = torch.cat([_draw(r, counts[r], n, 0.9, 1e-3) for r in range(world)]).double()
    ref = expected.var(0, unbiased=True)
    for _, samples, mean, var in results:
        assert torch.allclose(mean.double(), expected.mean(0), atol=1e-6)
        assert ((var.double() - ref).abs() / ref).max() < 1e-3
    one_pass = (expected.float().square().sum(0) - 256 * expected.float().mean(0) ** 2) / 255
    assert ((one_pass.double() - ref).abs() / ref).max() > 0.05   # what the previous formula delivered


def test_combine_posterior_single_process():
    local = torch.randn(5, 12, generator=torch.Generator().manual_seed(0))
    gathered = local.clone()
    out = combine_posterior(gathered, gathered, local.sum(0), local.square().sum(0), [5])
    assert torch.equal(out.samples, local)
    assert torch.allclose(out.variance, local.var(0), atol=1e-6)


# ---------------------------------------------------------------- PSLD / ReSample: global-batch norms over ranks
def _psld_like_loss(z, y, dec_w, enc_w, norm):
    """A PSLD-shaped scalar: omega * ||y - A D(z)|| + gamma * ||z - E(x_eff)|| with linear stand-ins for D, E, A."""
    x0 = torch.tanh(z @ dec_w)
    r = y - 0.5 * x0
    lik = norm(r)
    x_eff = x0 + 0.5 * r
    glue = norm(z - torch.tanh(x_eff @ enc_w))
    return 0.1 * lik + 1.0 * glue


def _norm_worker(rank, world, port, q):
    from samplers_b200.distributed import AllReduceSum, global_norm, group_size, reduce_sum_
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(7)
        z_all, y_all = torch.randn(6, 8, generator=g), torch.randn(6, 12, generator=g)
        dec_w, enc_w = torch.randn(8, 12, generator=g) * 0.3, torch.randn(12, 8, generator=g) * 0.3
        lo, hi = rank * 3, rank * 3 + 3
        z = z_all[lo:hi].clone().requires_grad_()
        loss = _psld_like_loss(z, y_all[lo:hi], dec_w, enc_w, lambda t: global_norm(t, dist.group.WORLD))
        (grad,) = torch.autograd.grad(loss, z)
        s = reduce_sum_(torch.tensor(float(rank + 1)), dist.group.WORLD)
        a = AllReduceSum.apply(torch.tensor(2.0, requires_grad=True) * (rank + 1), dist.group.WORLD)
        q.put((rank, float(loss), grad, float(s), float(a), group_size(dist.group.WORLD), group_size(None)))
    finally:
        dist.destroy_process_group()


def test_global_norms_over_two_ranks_equal_the_full_batch():
    """Sharding a batch over 2 ranks with all-reduced sums of squares reproduces the reference's batch-global
    norms (psld.py:130,138; resample_kernels.py:27): same loss on both ranks, gradient = slice of the full one."""
    world = 2
    results = _run_world(_norm_worker, world, lambda q: (q,))
    g = torch.Generator().manual_seed(7)
    z_all, y_all = torch.randn(6, 8, generator=g), torch.randn(6, 12, generator=g)
    dec_w, enc_w = torch.randn(8, 12, generator=g) * 0.3, torch.randn(12, 8, generator=g) * 0.3
    z = z_all.clone().requires_grad_()
    full = _psld_like_loss(z, y_all, dec_w, enc_w, torch.norm)
    (full_grad,) = torch.autograd.grad(full, z)
    for rank, loss, grad, s, a, gs, gs_none in results:
        assert abs(loss - float(full)) < 1e-5 * abs(float(full))
        assert torch.allclose(grad, full_grad[rank * 3: rank * 3 + 3], rtol=1e-5, atol=1e-7)
        assert s == 3.0 and a == 6.0 and gs == 2 and gs_none == 1


def test_global_norm_single_process_is_torch_norm():
    from samplers_b200.distributed import global_norm, group_size, reduce_sum_
    x = torch.randn(4, 5, generator=torch.Generator().manual_seed(1))
    assert torch.equal(global_norm(x), torch.norm(x)) and group_size(None) == 1
    t = torch.tensor(3.0)
    assert reduce_sum_(t, None) is t and float(t) == 3.0
