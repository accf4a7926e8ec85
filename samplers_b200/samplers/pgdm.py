"""PGDMSampler -- Pseudoinverse-Guided Diffusion Models (Song et al., 2023) with the call signature of
the reference sampler (samplers/samplers/pgdm.py:17-147).  For the operators that have an exact
pseudo-inverse of the form A^+ = c A^T (identity, inpainting: c = 1; f x f box: c = f^2) the consistency
loss || A^+ y - A^+ A x0 ||^2 (pgdm.py:116-121) has gradient -2c A^T (y - A x0) with respect to x0, so the
step is the DPS pair of kernels with likelihood weight 2c and a fixed guidance scale

    x_{t-1} = bridge(x_t, x0) + guidance_weight * sqrt(1 - acp_t) * (cot - s1 * VJP(cot))     (pgdm.py:130-135)

Operators without ``apply_pseudo_inverse`` raise NotImplementedError, as in the reference (pgdm.py:55-66).
"""
from __future__ import annotations

from typing import Generic, TypeVar

import torch
from torch import Tensor

from ..inverse_problem import InverseProblem
from .base import PosteriorSampler
from .dps import DPSRun, _default_draw
from .utils.batch_view import BatchView

Condition_co = TypeVar("Condition_co", covariant=True)


class PGDMSampler(PosteriorSampler, Generic[Condition_co]):
    draw = staticmethod(_default_draw)

    def __call__(self, inverse_problem: InverseProblem, num_sampling_steps: int = 50, num_reconstructions: int = 1,
                 guidance_weight: float = 1.0, eta: float = 1.0, condition: Condition_co | None = None,
                 keep_reconstruction_dim: bool = False, *args, **kwargs) -> Tensor:
        operator = inverse_problem.operator
        try:
            gain = float(operator._pinv_gain())
        except NotImplementedError as exc:
            raise NotImplementedError(
                "The operator in the inverse_problem must implement 'apply_pseudo_inverse' for PGDM.") from exc
        if args or kwargs:
            print(f"Warning: Unused args={args}, kwargs={kwargs} in PGDMSampler")
        view = BatchView(batch_shape=inverse_problem.batch_shape, num_samples=num_reconstructions,
                         data_shape=operator.x_shape)
        net = self._epsilon_network
        net.set_sampling_parameters(num_sampling_steps=num_sampling_steps, num_reconstructions=num_reconstructions,
                                    batch_size=view.batch_size)
        net.set_condition(condition=condition)
        try:
            gw = torch.tensor(float(guidance_weight), dtype=torch.float32)

            def scale(sc):  # guidance_weight * sqrt(1 - acp_t) in fp32, as pgdm.py:132-134
                return float(gw * torch.tensor(sc.sqrt_1m_acp, dtype=torch.float32))

            run = DPSRun(net, inverse_problem, view, 0.0, eta, self.draw, weight=2.0 * gain, fixed_scale=scale,
                         philox_seed=self.philox_seed, state_dtype=self.state_dtype)
            if self.cuda_graph:
                run.capture()
            for k in range(run.num_steps):
                run.step(k)
            x0 = view.unflatten(run.finalize().view(view.flat_shape))
        finally:
            net.clear_condition()
            net.clear_sampling_parameters()
        if num_reconstructions == 1 and not keep_reconstruction_dim:
            x0 = x0.squeeze(len(view.batch_shape))
        return x0
