from .base import PosteriorSampler
from .dps import DPSRun, DPSSampler

__all__ = ["PosteriorSampler", "DPSSampler", "DPSRun"]
