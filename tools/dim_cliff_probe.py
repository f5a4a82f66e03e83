#!/usr/bin/env python3
"""Throughput by dimension for the sampler x target pairs that have exact-dimension kernels plus a generic fallback:
shows where a dimension falls off a cliff.  Usage: python tools/dim_cliff_probe.py [rwros] [mala] [ram] [rwgp]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import amh_b200 as amh   # noqa: E402
from bench_configs import timed, spd   # noqa: E402

which = sys.argv[1:] or ["rwros", "rwgp", "mala", "ram"]
eng = amh.default_engine(0)
seeds = lambda n, s: np.random.default_rng(s).integers(0, 2 ** 64, size=n, dtype=np.uint64)


def go(tag, t, s, d, n, nsteps, spl, warmup=False, init=None):
    run = eng.run(eng.target(t.kind, d, t.blob()), s.lower(eng, d), n, seeds(n, 1), init)
    ms = timed(run, nsteps, warmup=warmup, spl=spl)
    st = run.state()
    print(f"{tag:28s} d={d:3d}  {n * nsteps / (ms * 1e-3):10.4g} chain-steps/s   accept={st['naccept'].sum() / (n * st['step']):.3f}", flush=True)
    run.close()


if "rwros" in which:
    for d in (5, 6, 7, 8, 9, 10, 11, 12, 14, 16, 20, 24, 28, 32):
        go("RWMH x Rosenbrock", amh.RosenbrockTarget(d), amh.RWMH(amh.MvNormal(np.zeros(d), (0.3 ** 2) * amh.I)), d, 65536, 200, 100)
if "rwgp" in which:
    for d in (5, 7, 8, 9, 12, 14, 16, 24, 28, 32):
        Sg = spd(d, 32, 1.0, 100.0)
        go("RWMH x GaussianPrecision", amh.GaussianPrecisionTarget(np.linalg.inv(Sg)), amh.RWMH(amh.MvNormal(np.zeros(d), (2.38 ** 2 / d) * Sg)), d, 65536, 200, 100)
if "static" in which:
    # StaticMH(MvNormal) as the reference constructs it: issymmetric = false, the Hastings term needs logq(state) and logq(cand)
    for d in (2, 4, 8, 10, 16, 24, 32):
        Sg = spd(d, 32, 1.0, 100.0)
        go("StaticMH x MvNormal", amh.MvNormalTarget(None, Sg), amh.StaticMH(amh.MvNormal(np.zeros(d), 1.3 * Sg)), d, 65536, 200, 100)
    for d in (2, 8, 16, 32):
        Sg = spd(d, 32, 1.0, 100.0)
        go("RWMH nonzero-mean proposal", amh.MvNormalTarget(None, Sg), amh.RWMH(amh.MvNormal(0.01 * np.ones(d), (2.38 ** 2 / d) * Sg)), d, 65536, 200, 100)
if "logistic" in which:
    # RWMH / MALA / RAM on the logistic-regression target (BASELINE config 4's model): only MALA has a dedicated kernel (K3L)
    for d, nrows, n in ((128, 10000, 16384), (32, 2000, 16384), (8, 500, 65536)):
        rng = np.random.default_rng(128)
        X = rng.normal(size=(nrows, d)) / np.sqrt(d)
        y = (rng.random(nrows) < 1 / (1 + np.exp(-X @ rng.normal(size=d)))).astype(float)
        t = amh.LogisticRegressionTarget(X, y, tau=10.0)
        go(f"RWMH x logistic {nrows} rows", t, amh.RWMH(amh.MvNormal(np.zeros(d), (0.05 ** 2) * amh.I)), d, n, 8, 4, init=np.zeros((d, n)))
        sg2 = 0.002
        go(f"MALA x logistic {nrows} rows", t, amh.MALA(lambda g: amh.MvNormal(0.5 * sg2 * g, sg2 * amh.I)), d, n, 8, 4, init=np.zeros((d, n)))
if "mala" in which:
    for d in (4, 5, 6, 7, 8, 9, 10, 12, 14, 16, 20, 24, 32):
        Sg = spd(d, 32, 0.5, 2.0)
        go("MALA x MvNormal", amh.MvNormalTarget(None, Sg), amh.MALA((lambda sg2: (lambda g: amh.MvNormal(0.5 * sg2 * g, sg2 * amh.I)))((0.3 / d ** (1 / 3)) ** 2)), d, 65536, 100, 50, init=np.zeros((d, 65536)))
if "ram" in which:
    for d in (2, 4, 8, 12, 15, 16, 17, 24, 32, 48, 64):
        Sg = spd(d, 64, 1e-2, 1.0)
        go("RAM warm-up x MvNormal", amh.MvNormalTarget(None, Sg), amh.RobustAdaptiveMetropolis(), d, 16384, 64, 16, warmup=True, init=np.zeros((d, 16384)))
