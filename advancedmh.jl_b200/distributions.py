"""The small slice of Distributions.jl the reference's proposal constructors use
(`Normal`, `MvNormal`; test/runtests.jl:58-60,78-80, README.md:35,104-112; `InverseGamma` in arrays of
univariate proposals, README.md:106, test/emcee.jl:19): plain parameter holders that the sampler layer lowers to
an `amh_sampler_desc`.  Univariate laws carry `component()` = (family, p0, p1, logc) in the parametrisation of
include/amh_contract.h (AMH_FAM_*)."""
from __future__ import annotations

import math

import numpy as np

FAM_NORMAL, FAM_INVGAMMA, FAM_GAMMA, FAM_UNIFORM, FAM_EXPONENTIAL, FAM_LOGNORMAL = 1, 2, 3, 4, 5, 6
_HALF_LOG_2PI = 0.5 * math.log(2.0 * math.pi)


class _Identity:
    """LinearAlgebra.I"""
    def __rmul__(self, c):
        return _ScaledIdentity(float(c))
    def __mul__(self, c):
        return _ScaledIdentity(float(c))


class _ScaledIdentity:
    def __init__(self, c):
        self.c = c
    def __rmul__(self, c):
        return _ScaledIdentity(self.c * float(c))
    __mul__ = __rmul__


I = _Identity()


class Normal:
    """Normal(mu, sigma)"""
    def __init__(self, mu=0.0, sigma=1.0):
        if not sigma > 0:
            raise ValueError("Normal: sigma must be positive")
        self.mu, self.sigma = float(mu), float(sigma)
    def __len__(self):
        return 1
    def component(self):
        return FAM_NORMAL, self.mu, self.sigma, -math.log(self.sigma) - _HALF_LOG_2PI


class LogNormal:
    """LogNormal(mu, sigma)"""
    def __init__(self, mu=0.0, sigma=1.0):
        if not sigma > 0:
            raise ValueError("LogNormal: sigma must be positive")
        self.mu, self.sigma = float(mu), float(sigma)
    def __len__(self):
        return 1
    def component(self):
        return FAM_LOGNORMAL, self.mu, self.sigma, -math.log(self.sigma) - _HALF_LOG_2PI


class InverseGamma:
    """InverseGamma(shape, scale)  (README.md:106 `InverseGamma(2,3)`)"""
    def __init__(self, shape=1.0, scale=1.0):
        if not (shape > 0 and scale > 0):
            raise ValueError("InverseGamma: shape and scale must be positive")
        self.shape, self.scale = float(shape), float(scale)
    def __len__(self):
        return 1
    def component(self):
        return FAM_INVGAMMA, self.shape, self.scale, self.shape * math.log(self.scale) - math.lgamma(self.shape)


class Gamma:
    """Gamma(shape, scale)"""
    def __init__(self, shape=1.0, scale=1.0):
        if not (shape > 0 and scale > 0):
            raise ValueError("Gamma: shape and scale must be positive")
        self.shape, self.scale = float(shape), float(scale)
    def __len__(self):
        return 1
    def component(self):
        return FAM_GAMMA, self.shape, self.scale, -self.shape * math.log(self.scale) - math.lgamma(self.shape)


class Uniform:
    """Uniform(a, b)"""
    def __init__(self, a=0.0, b=1.0):
        if not b > a:
            raise ValueError("Uniform: need a < b")
        self.a, self.b = float(a), float(b)
    def __len__(self):
        return 1
    def component(self):
        return FAM_UNIFORM, self.a, self.b, -math.log(self.b - self.a)


class Exponential:
    """Exponential(scale)"""
    def __init__(self, scale=1.0):
        if not scale > 0:
            raise ValueError("Exponential: scale must be positive")
        self.scale = float(scale)
    def __len__(self):
        return 1
    def component(self):
        return FAM_EXPONENTIAL, self.scale, 0.0, -math.log(self.scale)


UNIVARIATE = (Normal, LogNormal, InverseGamma, Gamma, Uniform, Exponential)


class MvNormal:
    """MvNormal(mu, Sigma).  Sigma: `I`, `c * I`, a vector of variances is NOT accepted
    (as in Distributions >= 0.25); a 1-d array is a diagonal of VARIANCES via Diagonal;
    a 2-d array is a full covariance (Cholesky-factorised here like PDMats does)."""
    def __init__(self, mu, Sigma=None):
        if Sigma is None:            # MvNormal(Sigma) form: zero mean
            Sigma, mu = mu, None
        if isinstance(Sigma, _Identity):
            Sigma = _ScaledIdentity(1.0)
        if isinstance(Sigma, _ScaledIdentity):
            if mu is None:
                raise ValueError("MvNormal(c*I) needs a mean to fix the dimension")
            self.dim = len(mu)
            self.kind = "scalar"
            if not Sigma.c > 0:
                raise ValueError("MvNormal: covariance must be positive definite")
            self.var = float(Sigma.c)
            self.scale = np.array([np.sqrt(Sigma.c)])
        else:
            S = np.asarray(Sigma, dtype=np.float64)
            if S.ndim == 1:
                self.dim = S.size
                self.kind = "diag"
                if not np.all(S > 0):
                    raise ValueError("MvNormal: covariance must be positive definite")
                self.scale = np.sqrt(S)
            else:
                self.dim = S.shape[0]
                self.kind = "full"
                L = np.linalg.cholesky(S)
                self.scale = L[np.tril_indices(self.dim)]      # packed by rows
                self.L = L
        if mu is None:
            mu = np.zeros(self.dim)
        mu = np.asarray(mu, dtype=np.float64)
        if mu.size != self.dim:
            raise ValueError("MvNormal: mean and covariance dimensions differ")
        self.mu = mu
        self.zero_mean = bool(np.all(mu == 0.0))
    def __len__(self):
        return self.dim

    @classmethod
    def from_cholesky(cls, mu, L):
        """full covariance given directly by its lower Cholesky factor"""
        L = np.asarray(L, dtype=np.float64)
        obj = cls.__new__(cls)
        obj.dim = L.shape[0]
        obj.kind = "full"
        obj.L = np.tril(L)
        obj.scale = obj.L[np.tril_indices(obj.dim)]
        mu = np.zeros(obj.dim) if mu is None else np.asarray(mu, dtype=np.float64)
        obj.mu = mu
        obj.zero_mean = bool(np.all(mu == 0.0))
        return obj


def Zeros(d):
    """FillArrays.Zeros(d)"""
    return np.zeros(int(d))
