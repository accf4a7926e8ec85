"""PosteriorSampler base (API of samplers/samplers/base.py:7-25)."""
from abc import ABC

from ..dtypes import Shape, Tensor
from ..networks import EpsilonNetwork


class PosteriorSampler(ABC):
    def __init__(self, network: EpsilonNetwork, cuda_graph: bool = False, process_group=None,
                 philox_seed: int | None = None, state_dtype=None):
        """``network`` as in the reference.  ``cuda_graph=True`` (DPS / PGDM) records one guided timestep as a CUDA
        graph and replays it for the whole loop; it needs a network whose ``forward`` takes the timestep as a
        device tensor without synchronising, and fails loudly otherwise (there is no silent eager fallback).
        ``process_group`` (PSLD / ReSample): see samplers_b200/distributed.py.  ``philox_seed`` (DPS / PGDM): draw
        the per-step noise inside the update kernel (Philox keyed by the seed) instead of with ``torch.randn``;
        reproducible across eager / graph execution, not bit-comparable with torch's stream."""
        self._epsilon_network = network
        self.cuda_graph = bool(cuda_graph)
        #: PSLD / ReSample: ranks over which the batch-global norms are all-reduced (None: rank-local norms)
        self.process_group = process_group
        self.philox_seed = philox_seed
        #: DPS / PGDM: torch.bfloat16 stores the sampler state, eps, cotangent and VJP as bf16 (18 instead of 40
        #: B/element per step; identity / mask / 4x box / separable-blur operators); default float32
        import torch
        self.state_dtype = torch.float32 if state_dtype is None else state_dtype

    @staticmethod
    def _flatten_leading(x: Tensor, *, x_shape: Shape) -> tuple[Tensor, Shape]:
        lead = x.shape[: x.ndim - len(x_shape)]
        return x.reshape(-1, *x_shape), lead

    @staticmethod
    def _unflatten_leading(x_flat: Tensor, *, batch_shape: tuple[int, ...]) -> Tensor:
        return x_flat.reshape(*batch_shape, *x_flat.shape[1:])
