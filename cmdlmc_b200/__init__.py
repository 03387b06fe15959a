"""cmdlmc_b200 -- B200-native (sm_100a) implementation of the cMD/LMC per-frame hot path:
O-O minimum-image distances -> cutoff neighbour lists -> jump rates -> KMC/LMC sweep.

Host classes keep the reference's names and signatures (gkabbe/cMDLMC, package `mdlmc`); all
arithmetic runs in hand-written CUDA kernels behind the C ABI of include/cmdlmc_b200.h.
"""
from .atombox import (AtomBox, AtomBoxCubic, AtomBoxMonoclinic, AtomBoxWater,  # noqa: F401
                      AtomBoxWaterLinearConversion, AtomBoxWaterRampConversion)
from .jumprate import ActivationEnergy, Exponential, Fermi, FermiAngle, JumpRate  # noqa: F401

__version__ = "0.1.0"
