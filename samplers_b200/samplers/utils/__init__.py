from .batch_view import BatchView
from .bridge_kernels import (BridgeStatistics, StepScalars, bridge_coefficients,
                             compute_bridge_kernel_statistics, ddim_step, plan_steps, sample_bridge_kernel)

__all__ = ["BatchView", "BridgeStatistics", "StepScalars", "bridge_coefficients", "plan_steps",
           "compute_bridge_kernel_statistics", "sample_bridge_kernel", "ddim_step"]
