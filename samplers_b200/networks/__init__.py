from .base import EpsilonNetwork, LatentEpsilonNetwork, NoCondition
from .ddpm import DDPMNetwork, OfflineDDPMPipeline
from .schedulers import DDPMSchedulerLite
from .sd15 import OfflineSDPipeline, StableDiffusionCondition, StableDiffusionNetwork
from .unet2d import UNet2DModel

__all__ = ["EpsilonNetwork", "LatentEpsilonNetwork", "NoCondition", "DDPMNetwork", "OfflineDDPMPipeline",
           "DDPMSchedulerLite", "UNet2DModel", "StableDiffusionNetwork", "StableDiffusionCondition", "OfflineSDPipeline"]
