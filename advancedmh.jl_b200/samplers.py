"""Sampler constructors with the reference's names and argument meaning:

    StaticProposal / RandomWalkProposal / Symmetric*      src/proposal.jl:1-21
    MetropolisHastings, StaticMH, RWMH                    src/mh-core.jl:44-51
    Ensemble, StretchProposal                             src/emcee.jl:1-4,63-68
    MALA                                                  src/MALA.jl:1-11
    RobustAdaptiveMetropolis                              src/RobustAdaptiveMetropolis.jl:75-87

Each lowers to one POD `amh_sampler_desc`; anything the device cannot represent
raises ValueError (ArgumentError in the Julia shim)."""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

from . import _capi as K
from .distributions import MvNormal, Normal, UNIVARIATE


# ------------------------------------------------------------------ proposals
class Proposal:
    issymmetric = False
    def __init__(self, proposal, issymmetric=None):
        self.proposal = proposal
        if issymmetric is not None:
            self.issymmetric = bool(issymmetric)


class StaticProposal(Proposal):
    """StaticProposal{issymmetric}(proposal); default issymmetric=false (proposal.jl:8)"""


class RandomWalkProposal(Proposal):
    """RandomWalkProposal{issymmetric}(proposal); default issymmetric=false (proposal.jl:18)"""


def SymmetricStaticProposal(p):
    return StaticProposal(p, True)


def SymmetricRandomWalkProposal(p):
    return RandomWalkProposal(p, True)


def _is_univariate_array(p):
    return isinstance(p, (list, tuple)) and len(p) > 0 and all(isinstance(q, UNIVARIATE) for q in p)


def _lower_gaussian(p):
    """Distribution | list of Normal -> (dim, cov_kind, mean|None, scale)"""
    if isinstance(p, MvNormal):
        kind = {"scalar": K.COV_SCALAR, "diag": K.COV_DIAG, "full": K.COV_FULL}[p.kind]
        return p.dim, kind, (None if p.zero_mean else p.mu), p.scale
    if isinstance(p, Normal):
        return 1, K.COV_SCALAR, (None if p.mu == 0.0 else np.array([p.mu])), np.array([p.sigma])
    if isinstance(p, (list, tuple)) and p and all(isinstance(q, Normal) for q in p):
        mu = np.array([q.mu for q in p])
        return len(p), K.COV_DIAG, (None if np.all(mu == 0) else mu), np.array([q.sigma for q in p])
    if callable(p):
        raise ValueError("function-valued proposals (proposal.jl:92-126) are host-only closures; "
                         "the device path supports Normal / MvNormal / arrays of Normal")
    raise ValueError(f"unsupported proposal distribution for the device path: {type(p).__name__}")


class MHSampler:
    pass


class MetropolisHastings(MHSampler):
    """MetropolisHastings(proposal)  (mh-core.jl:44-46).  `proposal` is

    * one StaticProposal / RandomWalkProposal over a Normal, an MvNormal, or an ARRAY of univariate distributions
      (`StaticProposal([Normal(0,1), InverseGamma(2,3)])`, README.md:106; proposal.jl:26-35), or
    * an ARRAY (or NamedTuple-like dict, in field order) of such proposals over univariate distributions, one per
      coordinate, each static or random-walk with its own `issymmetric` (README.md:111,125-133; proposal.jl:132-175,
      199-240).  The device state is a flat vector, so a NamedTuple only contributes its field order and names."""
    def __init__(self, proposal):
        self.components = None
        self.names = None
        if isinstance(proposal, dict):
            self.names = list(proposal.keys())
            proposal = list(proposal.values())
        if isinstance(proposal, (list, tuple)):
            if not proposal or not all(isinstance(q, (StaticProposal, RandomWalkProposal)) and isinstance(q.proposal, UNIVARIATE)
                                       for q in proposal):
                raise ValueError("an array / NamedTuple of proposals must hold Static/RandomWalkProposal objects over "
                                 "univariate distributions, one per coordinate")
            self.proposal = list(proposal)
            self.dim = len(proposal)
            self.components = [q.proposal.component() + (isinstance(q, RandomWalkProposal), q.issymmetric) for q in proposal]
            self.kind = K.SAMPLER_MIXED
            return
        if not isinstance(proposal, (StaticProposal, RandomWalkProposal)):
            raise ValueError("MetropolisHastings needs a StaticProposal or RandomWalkProposal (or an array of them)")
        self.proposal = proposal
        self.kind = K.SAMPLER_RW if isinstance(proposal, RandomWalkProposal) else K.SAMPLER_STATIC
        p = proposal.proposal
        if _is_univariate_array(p) and not all(isinstance(q, Normal) for q in p):
            self.dim = len(p)
            self.components = [q.component() for q in p]
        elif isinstance(p, UNIVARIATE) and not isinstance(p, Normal):
            self.dim = 1
            self.components = [p.component()]
        else:
            self.dim, self.cov_kind, self.mean, self.scale = _lower_gaussian(p)

    def lower(self, eng, dim):
        if dim != self.dim:
            raise ValueError(f"proposal dimension {self.dim} != model dimension {dim}")
        if self.components is not None:
            sym = False if self.kind == K.SAMPLER_MIXED else self.proposal.issymmetric
            return eng.sampler(kind=self.kind, dim=dim, symmetric=sym, cov_kind=K.COV_COMPONENTS, components=self.components)
        return eng.sampler(kind=self.kind, dim=dim, symmetric=self.proposal.issymmetric, cov_kind=self.cov_kind,
                           mean=self.mean, scale=self.scale)


def StaticMH(d):
    """StaticMH(d) / StaticMH(d::Int) = MvNormal(Zeros(d), I)  (mh-core.jl:48-49)"""
    if isinstance(d, (int, np.integer)):
        from .distributions import I
        d = MvNormal(np.zeros(int(d)), I)
    return MetropolisHastings(StaticProposal(d))


def RWMH(d):
    """RWMH(d) / RWMH(d::Int)  (mh-core.jl:50-51)"""
    if isinstance(d, (int, np.integer)):
        from .distributions import I
        d = MvNormal(np.zeros(int(d)), I)
    return MetropolisHastings(RandomWalkProposal(d))


# -------------------------------------------------------------------- emcee
class StretchProposal(Proposal):
    """StretchProposal(p, a=2.0)  (emcee.jl:63-68); `p` is only the law of the initial draw"""
    def __init__(self, proposal, stretch_length=2.0):
        super().__init__(proposal)
        self.stretch_length = float(stretch_length)


class Ensemble(MHSampler):
    """Ensemble(n_walkers, proposal)  (emcee.jl:1-4)"""
    def __init__(self, n_walkers, proposal):
        if not isinstance(proposal, StretchProposal):
            raise ValueError("Ensemble needs a StretchProposal")
        self.n_walkers, self.proposal = int(n_walkers), proposal

    def lower(self, eng, dim):
        p = self.proposal.proposal
        cov_kind, mean, scale = K.COV_SCALAR, None, None
        if _is_univariate_array(p) and not all(isinstance(q, Normal) for q in p):
            # StretchProposal([InverseGamma(2,3), Normal(0,1)]) (test/emcee.jl:19): the law of the initial draw
            if len(p) != dim:
                raise ValueError(f"proposal dimension {len(p)} != model dimension {dim}")
            return eng.sampler(kind=K.SAMPLER_STRETCH, dim=dim, cov_kind=K.COV_COMPONENTS,
                               components=[q.component() for q in p],
                               stretch_a=self.proposal.stretch_length, n_walkers=self.n_walkers)
        if p is not None:
            try:
                pdim, cov_kind, mean, scale = _lower_gaussian(p)
            except ValueError:
                pdim, scale = dim, None        # initial draw must then come from initial_params
            if pdim != dim:
                raise ValueError(f"proposal dimension {pdim} != model dimension {dim}")
        return eng.sampler(kind=K.SAMPLER_STRETCH, dim=dim, cov_kind=cov_kind, mean=mean, scale=scale,
                           stretch_a=self.proposal.stretch_length, n_walkers=self.n_walkers)


# --------------------------------------------------------------------- MALA
class MALA(MHSampler):
    """MALA(g -> MvNormal(c .* g, sigma2 * I))  (MALA.jl:1-11, README.md:180).

    The closure is opaque; its SHAPE is recovered by probing on the host
    (SURVEY.md 7 hard part 2): p(0) gives sigma2, p(e_i) gives the drift
    coefficient, p(2 e_i) checks linearity.  Anything else -> ValueError."""
    def __init__(self, proposal):
        if isinstance(proposal, RandomWalkProposal):
            proposal = proposal.proposal
        if not callable(proposal):
            raise ValueError("MALA needs a function g -> MvNormal(c*g, sigma2*I)")
        self.proposal = proposal
        self._probed = {}

    def probe(self, dim):
        if dim in self._probed:
            return self._probed[dim]
        z = np.zeros(dim)
        p0 = self.proposal(z)
        if not isinstance(p0, MvNormal) or p0.kind != "scalar" or p0.dim != dim or not p0.zero_mean:
            raise ValueError("MALA on the device needs proposal(g) = MvNormal(c*g, sigma2*I)")
        sigma2 = p0.var
        c = None
        for i in range(dim):
            e = np.zeros(dim); e[i] = 1.0
            p1, p2 = self.proposal(e), self.proposal(2 * e)
            ci = p1.mu[i]
            off = np.delete(p1.mu, i)
            if (p1.kind != "scalar" or np.any(off != 0) or abs(p2.mu[i] - 2 * ci) > 1e-12 * max(1.0, abs(ci))
                    or abs(p1.scale[0] - p0.scale[0]) > 0):
                raise ValueError("MALA proposal is not of the form MvNormal(c*g, sigma2*I)")
            if c is None:
                c = ci
            elif ci != c:
                raise ValueError("MALA proposal drift coefficient must be the same for every coordinate")
        self._probed[dim] = (sigma2, float(c))
        return self._probed[dim]

    def lower(self, eng, dim):
        sigma2, c = self.probe(dim)
        return eng.sampler(kind=K.SAMPLER_MALA, dim=dim, mala_sigma2=sigma2, mala_drift=c)


# ---------------------------------------------------------------------- RAM
@dataclass
class RobustAdaptiveMetropolis(MHSampler):
    """RobustAdaptiveMetropolis(; alpha=0.234, gamma=0.6, S=nothing, eigenvalue_lower_bound=0,
    eigenvalue_upper_bound=Inf)  (RobustAdaptiveMetropolis.jl:75-87); the Greek field names of the
    reference are spelled out."""
    alpha: float = 0.234
    gamma: float = 0.6
    S: object = None
    eigenvalue_lower_bound: float = 0.0
    eigenvalue_upper_bound: float = math.inf

    def lower(self, eng, dim):
        S0 = None
        if self.S is not None:
            S0 = np.asarray(self.S, dtype=np.float64)
            if S0.shape != (dim, dim):
                # RobustAdaptiveMetropolis.jl:202-204
                raise ValueError("The provided `S` has the wrong dimensionality.")
        return eng.sampler(kind=K.SAMPLER_RAM, dim=dim, ram_alpha=self.alpha, ram_gamma=self.gamma,
                           ram_eig_lo=self.eigenvalue_lower_bound, ram_eig_hi=self.eigenvalue_upper_bound,
                           ram_S0=S0)
