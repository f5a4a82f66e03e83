#!/usr/bin/env python3
"""Shows that a threaded numpy GEMM right before a timed region slows the 2-GPU in-process sample() call down: the BLAS
worker threads spin on every core for a while after the product.  Usage: python tools/blas_spin_probe.py [ngpus]"""
import os, sys, time
import numpy as np
from threadpoolctl import threadpool_limits, threadpool_info
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import amh_b200 as amh   # noqa: E402
import bench             # noqa: E402
ngpus = int(sys.argv[1]) if len(sys.argv) > 1 else 2
d, per_gpu, spl = 32, 65536, 500
t, s, Sg = bench.make_problem(amh, d)
L = np.linalg.cholesky(Sg)
model = amh.DensityModel(t)
eng = amh.default_engine(0)
n = per_gpu * ngpus
hinit = eng.pinned_empty((d, n)); hinit[...] = L @ np.random.default_rng(100).normal(size=(d, n))
pout = eng.pinned_empty((2, d + 1, n)); pacc = eng.pinned_empty((2, n), dtype=np.uint8)
par = amh.MCMCB200(ngpus=ngpus)
print("BLAS:", [(i.get("internal_api"), i.get("num_threads")) for i in threadpool_info()], "cpus", os.cpu_count())


def calls(k):
    ts = []
    for i in range(k):
        t0 = time.perf_counter()
        amh.sample(model, s, par, 2, n, initial_params=hinit, thinning=spl, chain_type=amh.Chains, seed=i, out=(pout, pacc))
        ts.append(1e3 * (time.perf_counter() - t0))
    return " ".join(f"{x:6.2f}" for x in ts)


calls(3)
time.sleep(0.5)
print("quiet process                      ms per call:", calls(6))
z = np.random.default_rng(1).normal(size=(d, n))
y = L @ z
print("right after a threaded GEMM        ms per call:", calls(6))
with threadpool_limits(limits=1):
    y = L @ z
print("right after a single-threaded GEMM ms per call:", calls(6))
y = L @ z
time.sleep(0.5)
print("threaded GEMM, then 0.5 s of sleep ms per call:", calls(6))
