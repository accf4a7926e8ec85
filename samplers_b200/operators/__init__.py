from .base import LinearOperator, NonlinearOperator, Operator
from .blur import (GaussianBlurOperator, MotionBlurOperator, SeparableBlurOperator, gaussian_taps,
                   motion_line_kernel, motion_walk_kernel)
from .identity import IdentityOperator
from .inpainting import (CenterInpaintingOperator, CenterOutpaintingOperator, InpaintingOperator,
                         RandomInpaintingOperator, SidePaintingOperator, get_mask_inpaint_center,
                         get_mask_random, get_mask_side_painting)
from .superres import BoxDownsampleOperator, MaskedBoxDownsampleOperator, SuperResolutionOperator

__all__ = [
    "Operator", "NonlinearOperator", "LinearOperator", "IdentityOperator", "InpaintingOperator",
    "CenterInpaintingOperator", "CenterOutpaintingOperator", "SidePaintingOperator",
    "RandomInpaintingOperator", "GaussianBlurOperator", "SeparableBlurOperator", "MotionBlurOperator",
    "BoxDownsampleOperator", "MaskedBoxDownsampleOperator", "SuperResolutionOperator", "gaussian_taps", "motion_line_kernel",
    "motion_walk_kernel", "get_mask_inpaint_center", "get_mask_side_painting", "get_mask_random",
]
