#!/usr/bin/env python3
"""Which earlier stage of tools/job_check.py slows the in-process 2-GPU sample() call down (11.3 ms there, 7.0 ms in a
fresh process)?  Re-measures the call after every stage.  Usage: python tools/job_e2e_bisect.py [ngpus]"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import amh_b200 as amh   # noqa: E402
import bench             # noqa: E402
import job_check         # noqa: E402
ngpus = int(sys.argv[1]) if len(sys.argv) > 1 else 2
if "--torch-first" in sys.argv:
    import torch
    print("torch imported first; device_count", torch.cuda.device_count(), "threads", torch.get_num_threads(), flush=True)
d, per_gpu, spl = 32, 65536, 500
t, s, Sg = bench.make_problem(amh, d)
L = np.linalg.cholesky(Sg)
model = amh.DensityModel(t)
eng = amh.default_engine(0)
n = per_gpu * ngpus
hinit = eng.pinned_empty((d, n)); hinit[...] = L @ np.random.default_rng(100).normal(size=(d, n))
pout = eng.pinned_empty((2, d + 1, n)); pacc = eng.pinned_empty((2, n), dtype=np.uint8)


def measure(tag, reps=5):
    if "--fresh-job" in sys.argv:          # a NEW job (contexts, worker threads) for every measurement
        import advancedmh_jl_b200.sampling as S
        for j in S._JOBS.values():
            j.close()
        S._JOBS.clear()
    par = amh.MCMCB200(ngpus=ngpus)
    amh.sample(model, s, par, 2, n, initial_params=hinit, thinning=spl, chain_type=amh.Chains, seed=99, out=(pout, pacc))
    t0 = time.perf_counter()
    for i in range(reps):
        amh.sample(model, s, par, 2, n, initial_params=hinit, thinning=spl, chain_type=amh.Chains, seed=i, out=(pout, pacc))
    print(f"{tag:28s} {1e3 * (time.perf_counter() - t0) / reps:7.3f} ms per call", flush=True)


def k1_stage():
    # what tools/job_check.py does before its 2-GPU measurement: the same call on a 1-GPU job, buffers of its own, dropped
    n1 = per_gpu
    h1 = eng.pinned_empty((d, n1)); h1[...] = hinit[:, :n1]
    o1 = eng.pinned_empty((2, d + 1, n1)); a1 = eng.pinned_empty((2, n1), dtype=np.uint8)
    p1 = amh.MCMCB200(ngpus=1)
    for i in range(6):
        amh.sample(model, s, p1, 2, n1, initial_params=h1, thinning=spl, chain_type=amh.Chains, seed=i, out=(o1, a1))
    print("ran the 1-GPU job", flush=True)


if "--k1-first" in sys.argv:
    k1_stage()
measure("fresh process")
if "--only" in sys.argv:
    sys.exit(0)
out = {"ngpus": ngpus, "parity": {}, "broadcast": {}, "e2e": {}, "c5": {}}
orig = os.environ.get("AMH_JOB_BCAST")
for mode in ("h2d", "peer", "nccl"):
    os.environ["AMH_JOB_BCAST"] = mode
    job = eng.job(ngpus)
    tt = amh.MvNormalTarget(None, Sg)
    job.target(tt.kind, d, tt.blob())
    job.close()
    os.environ.pop("AMH_JOB_BCAST", None)
    measure(f"after a {mode} job")
job = eng.job(ngpus)
job_check.parity(eng, job, out)
measure("after parity")
job_check.c5(eng, job, ngpus, out)
measure("after c5 (job open)")
job.close()
measure("after c5 (job closed)")
if "--k1-late" in sys.argv:
    k1_stage()
    measure("after the 1-GPU job")
import torch          # after the library has loaded an NCCL of its own: must still import (AMH_NCCL_LIB, _capi.py)
print("torch imported after an NCCL job; NCCL", torch.cuda.nccl.version(), "device_count", torch.cuda.device_count(), flush=True)
measure("after import torch")
