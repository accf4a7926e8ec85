from .base import PosteriorSampler
from .dps import DPSRun, DPSSampler
from .psld import PSLDSampler

__all__ = ["PosteriorSampler", "DPSSampler", "PSLDSampler", "DPSRun"]
