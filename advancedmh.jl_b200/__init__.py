"""advancedmh.jl_b200 -- host-side mirror of the AdvancedMH.jl interface for the
multi-chain path, over the hand-written sm_100a engine in libamh_b200.so.

Importable as `advancedmh_jl_b200` through the loader `amh_b200.py` at the
repository root (the directory name contains a dot)."""
from . import _capi
from ._capi import AMHArgumentError, AMHError, AMHStateError, Engine, PosDefException, contract, precision
from .distributions import (Exponential, Gamma, I, InverseGamma, LogNormal, MvNormal, Normal, Uniform, Zeros)
from .models import (DensityModel, DeviceTarget, GaussianPrecisionTarget, IIDNormalTarget,
                     LogisticRegressionTarget, MvNormalTarget, NormalInverseGammaToy, RosenbrockTarget, SourceTarget)
from .samplers import (MALA, RWMH, Ensemble, MetropolisHastings, RandomWalkProposal,
                       RobustAdaptiveMetropolis, StaticMH, StaticProposal, StretchProposal,
                       SymmetricRandomWalkProposal, SymmetricStaticProposal)
from .sampling import (Chains, LocalParams, MCMCB200, MCMCDistributed, MCMCSerial, MCMCThreads, SamplerState, StructArray,
                       Transition, default_engine, sample, shard_bounds)

__all__ = [n for n in dir() if not n.startswith("_")]
