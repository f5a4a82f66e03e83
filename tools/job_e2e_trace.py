#!/usr/bin/env python3
"""Phase trace of the in-process multi-GPU sample() call (MCMCB200(ngpus=k)): AMH_TRACE prints the library's and the
Python layer's phase times on stderr.  Usage: AMH_TRACE=1 python tools/job_e2e_trace.py [ngpus] [calls]"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import amh_b200 as amh   # noqa: E402
import bench             # noqa: E402
ngpus = int(sys.argv[1]) if len(sys.argv) > 1 else 2
calls = int(sys.argv[2]) if len(sys.argv) > 2 else 3
d, per_gpu, spl = 32, 65536, 500
t, s, Sg = bench.make_problem(amh, d)
L = np.linalg.cholesky(Sg)
model = amh.DensityModel(t)
eng = amh.default_engine(0)
for k in (1, ngpus):
    n = per_gpu * k
    hinit = eng.pinned_empty((d, n)); hinit[...] = L @ np.random.default_rng(100).normal(size=(d, n))
    pout = eng.pinned_empty((2, d + 1, n)); pacc = eng.pinned_empty((2, n), dtype=np.uint8)
    par = amh.MCMCB200(ngpus=k)
    for i in range(calls):
        print(f"---- ngpus={k} call {i}", file=sys.stderr, flush=True)
        t0 = time.perf_counter()
        amh.sample(model, s, par, 2, n, initial_params=hinit, thinning=spl, chain_type=amh.Chains, seed=i, out=(pout, pacc))
        print(f"---- total {1e3 * (time.perf_counter() - t0):.3f} ms", file=sys.stderr, flush=True)
