"""B200-native DGMR generator / discriminator training step.

Drop-in module API of openclimatefix/skillful_nowcasting's `dgmr` package (dgmr/__init__.py:3-6):
same constructors, forward signatures, state-dict keys and `from_pretrained` contract; the arithmetic
runs as hand-written sm_100a CUDA reached through the C ABI in include/dgmr_b200.h.
There is no CPU / cuDNN / Triton fallback: without libdgmr_b200.so the forward raises.
"""
from .common import ContextConditioningStack, LatentConditioningStack
from .discriminators import Discriminator, SpatialDiscriminator, TemporalDiscriminator
from .generators import Generator, Sampler
from .dgmr import DGMR

__all__ = [
    "DGMR", "Generator", "Sampler", "ContextConditioningStack", "LatentConditioningStack",
    "Discriminator", "SpatialDiscriminator", "TemporalDiscriminator",
]
