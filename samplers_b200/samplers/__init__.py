from .base import PosteriorSampler
from .dps import DPSRun, DPSSampler
from .psld import PSLDSampler
from .resample import ReSampleSampler

__all__ = ["PosteriorSampler", "DPSSampler", "PSLDSampler", "ReSampleSampler", "DPSRun"]
