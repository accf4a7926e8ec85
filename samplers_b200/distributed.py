"""Multi-GPU posterior sampling: independent samples sharded across ranks, one
process per GPU, no collective inside the sampling loop (DPS samples are
independent: per-sample norm dps.py:118-120, per-sample log-prob noise.py:79).
NCCL is used only at the end for DPS:

  * all_gather of every rank's final x0 samples,
  * all_reduce(SUM) of the per-pixel [sum, sum of squares] -> posterior mean / variance.

The final Tweedie kernel (psx_tweedie) writes each rank's samples straight into
its slot of the gather buffer and produces the two moment buffers in the same pass.

PSLD and ReSample couple the samples of a batch through batch-global norms / means.  Sharded over
ranks they run either as replicas with rank-local norms (``process_group=None``, the default: each shard behaves
like a separate reference call) or with exact global-batch semantics (``process_group=<group>``): the sums of
squares behind every norm / MSE are all-reduced (``reduce_sum_`` / ``AllReduceSum`` / ``global_norm``), one
scalar per reduction, so the union of the shards reproduces one reference call on the concatenated batch.
"""
from __future__ import annotations

import dataclasses

import torch
import torch.distributed as dist
from torch import Tensor


def group_size(group) -> int:
    """Ranks sharing batch-global reductions: 1 when ``group`` is None or torch.distributed is not initialised."""
    if group is None or not dist.is_available() or not dist.is_initialized():
        return 1
    return dist.get_world_size(group)


def reduce_sum_(value: Tensor, group) -> Tensor:
    """In-place SUM all-reduce of a (scalar) tensor over ``group``; a no-op for a single rank.  Not recorded by
    autograd -- for use inside autograd.Function.forward."""
    if group_size(group) > 1:
        dist.all_reduce(value, op=dist.ReduceOp.SUM, group=group)
    return value


class AllReduceSum(torch.autograd.Function):
    """S = sum over ranks of s_r, differentiable: every rank holds the SAME scalar loss f(S) and wants its gradient
    with respect to its own shard, d f(S) / d s_r = f'(S), so the backward is the identity (no second reduction)."""

    @staticmethod
    def forward(ctx, value: Tensor, group):
        return reduce_sum_(value.detach().clone(), group)

    @staticmethod
    def backward(ctx, grad: Tensor):
        return grad, None


def global_norm(x: Tensor, group=None) -> Tensor:
    """Frobenius norm over the batch of ALL ranks of ``group`` (the reference's batch-global ``torch.norm`` when the
    batch is sharded: psld.py:130,138, resample_kernels.py:27).  With one rank it is ``torch.norm(x)`` itself."""
    if group_size(group) == 1:
        return torch.norm(x)
    return AllReduceSum.apply(x.square().sum(), group).sqrt()


def shard_count(total: int, rank: int, world: int) -> tuple[int, int]:
    """(start, count) of the contiguous block of ``total`` items owned by ``rank``."""
    if total < 0 or world < 1 or not (0 <= rank < world):
        raise ValueError("bad shard request")
    base, rem = divmod(total, world)
    count = base + (1 if rank < rem else 0)
    start = rank * base + min(rank, rem)
    return start, count


@dataclasses.dataclass
class PosteriorSummary:
    samples: Tensor   # (R_total, *x_shape) -- all ranks' reconstructions, rank-major
    mean: Tensor      # (*x_shape)
    variance: Tensor  # (*x_shape), unbiased (divides by R_total - 1)


def combine_posterior(local_slot: Tensor, gathered: Tensor, total: Tensor, total_sq: Tensor, counts: list[int],
                      group=None) -> PosteriorSummary:
    """Terminal exchange.  ``gathered`` is (world * max_count, n); this rank already wrote its samples into
    rows [rank * max_count, rank * max_count + counts[rank]) (``local_slot`` is that view).  ``total`` is this rank's
    per-pixel sum (from the final-estimate kernel) and is all-reduced to the posterior mean.  The variance is the
    two-pass form about that mean, sum((x - mean)^2) / (R - 1) over the gathered samples: the one-pass
    (sum x^2 - R mean^2) cancels catastrophically for a converged posterior (spread << mean), so ``total_sq`` is
    accepted for compatibility but not used."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    max_count = max(counts)
    if world > 1:
        dist.all_gather_into_tensor(gathered, gathered.view(world, max_count, -1)[dist.get_rank(group)].contiguous(),
                                    group=group)
        total = total.clone()
        dist.all_reduce(total, op=dist.ReduceOp.SUM, group=group)
    rows = [gathered.view(world, max_count, -1)[r, : counts[r]] for r in range(world)]
    samples = torch.cat(rows, dim=0)
    n_tot = float(sum(counts))
    mean = total / n_tot
    # every rank holds all samples after the gather: the second pass needs no further exchange
    var = (samples - mean).square_().sum(0) / max(n_tot - 1.0, 1.0)
    return PosteriorSummary(samples=samples, mean=mean, variance=var)


def sample_posterior(sampler, inverse_problem, *, num_reconstructions: int, num_sampling_steps: int = 50,
                     gamma: float = 1.0, eta: float = 1.0, condition=None, group=None) -> PosteriorSummary:
    """Draw ``num_reconstructions`` DPS samples of one observation across all ranks of ``group``
    (or on this GPU alone when torch.distributed is not initialised)."""
    if len(inverse_problem.batch_shape) != 0:
        raise ValueError("sample_posterior shards the reconstructions of a single observation")
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    counts = [shard_count(num_reconstructions, r, world)[1] for r in range(world)]
    if min(counts) < 1:
        raise ValueError("need at least one reconstruction per rank")
    run = sampler.prepare(inverse_problem, num_sampling_steps, counts[rank], gamma, eta, condition)
    try:
        if getattr(sampler, "cuda_graph", False):
            run.capture()
        for k in range(run.num_steps):
            run.step(k)
        max_count = max(counts)
        gathered = torch.zeros((world * max_count, run.n), device=run.device, dtype=torch.float32)
        slot = gathered[rank * max_count: rank * max_count + counts[rank]]
        total = torch.empty(run.n, device=run.device, dtype=torch.float32)
        total_sq = torch.empty(run.n, device=run.device, dtype=torch.float32)
        run.finalize(out=slot, total=total, total_sq=total_sq)
    finally:
        sampler.release()
    out = combine_posterior(slot, gathered, total, total_sq, counts, group)
    x_shape = tuple(inverse_problem.operator.x_shape)
    return PosteriorSummary(samples=out.samples.view(-1, *x_shape), mean=out.mean.view(x_shape),
                            variance=out.variance.view(x_shape))
