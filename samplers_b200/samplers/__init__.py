from .base import PosteriorSampler
from .dps import DPSRun, DPSSampler
from .pgdm import PGDMSampler
from .psld import PSLDSampler
from .resample import ReSampleSampler

__all__ = ["PosteriorSampler", "DPSSampler", "PSLDSampler", "ReSampleSampler", "PGDMSampler", "DPSRun"]
