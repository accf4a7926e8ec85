"""InverseProblem(operator, observation, noise) -- same dataclass surface as the
reference (samplers/inverse_problem.py:10-67): ``residual``, ``log_likelihood``,
``score``, ``batch_shape``, ``from_observation``, ``from_clean_data``."""
from __future__ import annotations

from dataclasses import dataclass

import torch

from .dtypes import RNG, Shape, Tensor
from .noise import NoiseModel
from .operators import Operator


@dataclass
class InverseProblem:
    operator: Operator
    observation: Tensor
    noise: NoiseModel

    def residual(self, x: Tensor) -> Tensor:
        return self.observation - self.operator(x)

    def log_likelihood(self, x: Tensor) -> Tensor:
        return self.noise.log_prob(self.residual(x))

    def score(self, x: Tensor) -> Tensor:
        return self.noise.score(self.residual(x)) * (-1)

    @property
    def batch_shape(self) -> Shape:
        return self.observation.shape[: self.observation.ndim - len(self.operator.y_shape)]

    @classmethod
    def from_observation(cls, obs: Tensor, *, operator: Operator, noise: NoiseModel) -> "InverseProblem":
        return cls(operator=operator, observation=obs, noise=noise)

    @classmethod
    def from_clean_data(cls, x_true: Tensor, *, operator: Operator, noise: NoiseModel,
                        rng: RNG = None) -> "InverseProblem":
        """Simulate y = A(x_true) + eps, eps ~ noise.sample (rng makes the draw reproducible).  The operator pass
        runs in its sm_100a kernel and the noise is folded in by psx_add_noise (same roundings as the reference's
        ``y_clean + eps``, inverse_problem.py:55-62); the random draw itself stays with torch's generator."""
        from . import _native
        with torch.no_grad():
            y = operator(x_true)
            if y.is_cuda and y.dtype == torch.float32:
                try:
                    raw, scale, shift = noise._draw_affine(y.shape, y.device, rng)
                except NotImplementedError:
                    raw = None
                if raw is not None:
                    y = y.contiguous()
                    if y.data_ptr() == x_true.data_ptr():      # identity operators may hand back a view of x_true
                        y = y.clone()
                    _native.add_noise(y, raw.contiguous(), scale, shift)
                    return cls(operator=operator, observation=y, noise=noise)
            y = y + noise.sample(shape=y.shape, device=y.device, dtype=y.dtype, generator=rng)
        return cls(operator=operator, observation=y, noise=noise)
