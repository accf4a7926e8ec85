"""samplers_b200 -- B200-native (sm_100a) implementation of the per-timestep
posterior-sampling update of thomashirtz/samplers, behind the reference's own
class API: ``InverseProblem`` / ``Operator`` / ``NoiseModel`` / ``Sampler.__call__``.

    from samplers_b200.networks import DDPMNetwork
    from samplers_b200.operators import GaussianBlurOperator
    from samplers_b200.noise import GaussianNoise
    from samplers_b200.inverse_problem import InverseProblem
    from samplers_b200.samplers import DPSSampler

The hot path lives in ``csrc/`` (CUDA) behind the C ABI of ``include/psx.h``;
the host side is Python + torch for device memory, streams and the eps-network.
"""
from .inverse_problem import InverseProblem
from .noise import GaussianNoise, NoiseModel, PoissonNoise

__version__ = "0.1.0"
__all__ = ["InverseProblem", "GaussianNoise", "PoissonNoise", "NoiseModel", "__version__"]
