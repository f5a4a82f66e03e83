"""Model side of the drop-in boundary.

The reference wraps an arbitrary Julia closure in `DensityModel(f)`
(src/AdvancedMH.jl:52-54) or takes any `LogDensityProblems` object
(src/AdvancedMH.jl:56,76).  A closure cannot run inside a CUDA kernel, so the
device path accepts models from a fixed catalogue; each
class below is the host-side description of one catalogue entry and lowers to
`amh_target_create(kind, dim, blob)`.  Anything else raises ValueError (the
Julia shim raises ArgumentError) -- there is no CPU fallback."""
from __future__ import annotations

import math
import numpy as np

from . import _capi as K


class DeviceTarget:
    kind: int
    dim: int
    def blob(self) -> np.ndarray:
        raise NotImplementedError
    # LogDensityProblems.dimension / capabilities
    def dimension(self):
        return self.dim
    has_gradient = False


class IIDNormalTarget(DeviceTarget):
    """theta = (mu, sigma);  sum(logpdf.(Normal(mu, sigma), data)) on sigma >= 0
    (README.md:26-31, test/runtests.jl:23-31)"""
    kind, dim, has_gradient = K.TARGET_IID_NORMAL, 2, True
    def __init__(self, data):
        self.data = np.asarray(data, dtype=np.float64).ravel()
    def blob(self):
        return self.data


class MvNormalTarget(DeviceTarget):
    """logpdf(MvNormal(mu, Sigma), x)  (test/RobustAdaptiveMetropolis.jl:1-9 `Gaussian`)"""
    kind, has_gradient = K.TARGET_MVNORMAL, True
    def __init__(self, mu, Sigma):
        Sigma = np.atleast_2d(np.asarray(Sigma, dtype=np.float64))
        self.dim = Sigma.shape[0]
        self.mu = np.zeros(self.dim) if mu is None else np.asarray(mu, dtype=np.float64).ravel()
        self.Sigma = Sigma
        C = np.linalg.cholesky(Sigma)
        self.U = np.linalg.inv(C)           # lower triangular, U'U = inv(Sigma)
        self.U = np.tril(self.U)
        logdet = 2.0 * np.sum(np.log(np.diag(C)))
        self.c0 = -0.5 * (self.dim * math.log(2 * math.pi) + logdet)
    def blob(self):
        return np.concatenate([[self.c0], self.mu, self.U[np.tril_indices(self.dim)]])


class GaussianPrecisionTarget(DeviceTarget):
    """-x'Ax/2 with gradient -Ax  (test/runtests.jl:335-347 `TheNormalLogDensity`)"""
    kind, has_gradient = K.TARGET_GAUSS_PREC, True
    def __init__(self, A):
        self.A = np.atleast_2d(np.asarray(A, dtype=np.float64))
        self.dim = self.A.shape[0]
    def blob(self):
        return self.A.ravel()


class RosenbrockTarget(DeviceTarget):
    """-sum_{i<d-1} [b (x_{i+1} - x_i^2)^2 + (a - x_i)^2] / s   (BASELINE config 3)"""
    kind, has_gradient = K.TARGET_ROSENBROCK, True
    def __init__(self, dim, a=1.0, b=100.0, scale=20.0):
        self.dim, self.a, self.b, self.s = int(dim), float(a), float(b), float(scale)
    def blob(self):
        return np.array([self.a, self.b, self.s])


class LogisticRegressionTarget(DeviceTarget):
    """Bayesian logistic regression with a N(0, tau^2 I) prior (BASELINE config 4)"""
    kind, has_gradient = K.TARGET_LOGISTIC, True
    def __init__(self, X, y, tau=10.0):
        self.X = np.ascontiguousarray(X, dtype=np.float64)
        self.y = np.asarray(y, dtype=np.float64).ravel()
        self.tau = float(tau)
        self.dim = self.X.shape[1]
    def blob(self):
        return np.concatenate([[self.tau], self.X.ravel(), self.y])


class NormalInverseGammaToy(DeviceTarget):
    """The emcee example of the reference's tests (test/emcee.jl:5-15 and, with
    log_space=True, :46-56): s ~ InverseGamma(alpha, beta), m ~ N(0, s), y_i ~ N(m, s)."""
    dim = 2
    def __init__(self, obs=(1.5, 2.0), alpha=2.0, beta=3.0, log_space=False):
        self.obs = np.asarray(obs, dtype=np.float64)
        self.alpha, self.beta = float(alpha), float(beta)
        self.kind = K.TARGET_NIG_TOY_LOG if log_space else K.TARGET_NIG_TOY
    def blob(self):
        cig = self.alpha * math.log(self.beta) - math.lgamma(self.alpha)
        return np.concatenate([[self.alpha, self.beta, cig], self.obs])


class SourceTarget(DeviceTarget):
    """A log-density stated as C++ source text (include/amh_user_target.h), compiled for the device by NVRTC
    (amh_target_create_source): the route from the catalogue to `DensityModel(f)` for an arbitrary `f`
    (src/AdvancedMH.jl:52-54) and, with `gradient=True`, to a LogDensityProblems object with
    `logdensity_and_gradient` (MALA.jl:100-105).  The source defines

        AMH_TARGET double amh_user_logdensity(const double* x, int dim, const double* data, long long ndata);
        AMH_TARGET void   amh_user_logdensity_and_gradient(const double* x, int dim, const double* data,
                                                           long long ndata, double* lp, double* grad);  // gradient=True

    `data` (optional float64 array) is copied to the device and passed to every call."""
    kind = K.TARGET_USER
    def __init__(self, dim, source, data=None, gradient=False):
        self.dim = int(dim)
        self.source = str(source)
        self.data = np.zeros(0) if data is None else np.ascontiguousarray(data, dtype=np.float64).ravel()
        self.has_gradient = bool(gradient)
    def blob(self):
        return self.data


class DensityModel:
    """DensityModel(target): same name and role as src/AdvancedMH.jl:52-54.  `target` is a catalogue entry or a
    `SourceTarget` (the closure stated as source text); a Python callable cannot run on the device."""
    def __init__(self, logdensity):
        if not isinstance(logdensity, DeviceTarget):
            raise ValueError(
                "DensityModel on the B200 path needs a catalogue target (IIDNormalTarget, MvNormalTarget, "
                "GaussianPrecisionTarget, RosenbrockTarget, LogisticRegressionTarget, NormalInverseGammaToy) or a "
                "SourceTarget (the log-density as source text); host closures cannot run on the device and there "
                "is no CPU fallback")
        self.logdensity = logdensity


def as_target(model) -> DeviceTarget:
    """accepts DensityModel(target) or a bare target (AbstractMCMC wraps LogDensityProblems
    objects in LogDensityModel the same way)"""
    if isinstance(model, DensityModel):
        return model.logdensity
    if isinstance(model, DeviceTarget):
        return model
    raise ValueError("model must be a DensityModel or a catalogue target")
