#!/usr/bin/env python3
"""The in-library multi-GPU job (amh_job_*, MCMCB200(ngpus=N)) on the GPUs of this box -- run under `gpurun --gpus N`:

  1. parity : N-device job == 1-device engine, bit for bit (RWMH C2 shape, stretch ensembles, RAM warm-up, MALA logistic)
  2. bcast  : the one-time broadcast of the C4 target blob (10.2 MB): nccl vs peer copies vs N host copies
  3. e2e    : sample(model, RWMH, MCMCB200(ngpus=N), ...) with host buffers, 65 536 chains per GPU x 500 steps,
              against the same call on 1 GPU (weak scaling of the single-process path a Julia host would use)
  4. c5     : BASELINE config 5 as stated: RAM d=64, 262 144 chains over 8 GPUs (32 768 per GPU), warm-up steps

Usage: python tools/job_check.py [ngpus]      (default: all visible devices)"""
import json
import os
import sys
import time

import numpy as np
from threadpoolctl import threadpool_limits

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import amh_b200 as amh   # noqa: E402
import bench             # noqa: E402


def seeds(n, s):
    return np.random.default_rng(s).integers(0, 2 ** 64, size=n, dtype=np.uint64)


def parity(eng, job, out):
    cases = []
    d = 32
    t, s, Sg = bench.make_problem(amh, d)
    cases.append(("rwmh_c2_shape", t, s, 8192 + 5, seeds(8192 + 5, 1), None, dict(), 60, False))
    dr, nw, ne = 10, 512, 5
    cases.append(("stretch_5x512", amh.RosenbrockTarget(dr), amh.Ensemble(nw, amh.StretchProposal(amh.MvNormal(np.zeros(dr), amh.I))),
                  nw * ne, seeds(ne, 2), None, dict(), 8, False))
    d5 = 64
    S5 = bench._spd(d5, 64, 1e-4, 1.0)
    cases.append(("ram_warmup_d64", amh.MvNormalTarget(None, S5), amh.RobustAdaptiveMetropolis(S=(2.38 / np.sqrt(d5)) * np.linalg.cholesky(S5)),
                  3000, seeds(3000, 3), np.zeros((d5, 3000)), dict(S=True), 24, True))
    rng = np.random.default_rng(128)
    X = rng.normal(size=(2000, 128)) / np.sqrt(128)
    y = (rng.random(2000) < 0.5).astype(float)
    cases.append(("mala_logistic_d128", amh.LogisticRegressionTarget(X, y, tau=10.0), amh.MALA(lambda g: amh.MvNormal((0.03 / 2) * g, 0.03 * amh.I)),
                  200, seeds(200, 4), np.zeros((128, 200)), dict(grad=True), 3, False))
    for name, t, s, n, sd, init, kw, steps, wu in cases:
        a = eng.run(eng.target_of(t), s.lower(eng, t.dim), n, sd, init)
        b = job.run(job.target_of(t), s.lower(job, t.dim), n, sd, init)
        a.steps(steps, warmup=wu); b.steps(steps, warmup=wu)
        sa, sb = a.state(**kw), b.state(**kw)
        same = all(np.array_equal(sa[k], sb[k], equal_nan=True) for k in sa if isinstance(sa[k], np.ndarray))
        oa, aa, _ = a.sample(3, 2, 2, 0, summary=False)
        ob, ab, _ = b.sample(3, 2, 2, 0, summary=False)
        same = same and np.array_equal(oa, ob) and np.array_equal(aa, ab)
        out["parity"][name] = {"bit_equal": bool(same), "shards": job.shards(), "broadcast": job.broadcast_info()[0]}
        print(f"parity {name:22s} {job.ngpus}-device job == 1 device: {same}", flush=True)
        a.close(); b.close()
        assert same, name


def bcast(eng, ngpus, out):
    d, nrows = 128, 10000
    rng = np.random.default_rng(128)
    X = rng.normal(size=(nrows, d)) / np.sqrt(d)
    y = (rng.random(nrows) < 0.5).astype(float)
    t = amh.LogisticRegressionTarget(X, y, tau=10.0)
    blob = t.blob()
    for mode in ("nccl", "peer", "h2d"):
        os.environ["AMH_JOB_BCAST"] = mode
        try:
            t0 = time.perf_counter()
            job = eng.job(ngpus)
            t_create = (time.perf_counter() - t0) * 1e3
            ms = []
            for _ in range(5):
                job.target(t.kind, d, blob)
                ms.append(job.broadcast_info()[1])
            got, _, init_ms = job.broadcast_info()
            out["broadcast"][mode] = {"mode_used": got, "ms_first": ms[0], "ms_best": min(ms), "bytes": int(blob.nbytes),
                                      "job_create_ms": t_create, "comm_init_ms": init_ms}
            print(f"bcast {mode:5s} ({got}): first {ms[0]:.2f} ms, best {min(ms):.2f} ms for {blob.nbytes / 1e6:.1f} MB to {ngpus} GPUs; job_create {t_create:.0f} ms (communicators {init_ms:.0f} ms)", flush=True)
            job.close()
        except amh.AMHError as e:
            out["broadcast"][mode] = {"error": str(e)}
            print(f"bcast {mode}: {e}", flush=True)
    os.environ.pop("AMH_JOB_BCAST", None)


def e2e(eng, ngpus, out, per_gpu=65536, spl=500, reps=5):
    d = 32
    t, s, Sg = bench.make_problem(amh, d)
    L = np.linalg.cholesky(Sg)
    model = amh.DensityModel(t)
    for k in sorted({1, ngpus}):
        n = per_gpu * k
        hinit = eng.pinned_empty((d, n))
        # one BLAS thread for the set-up product: numpy's OpenBLAS workers keep spinning on every core for ~100 ms after a
        # threaded GEMM, and a timed region that starts inside that window measures them, not the library (that was the
        # "0.57 efficiency at 2 GPUs" of the first version of this script: profiles/r2_job_fanout_2gpu.txt)
        with threadpool_limits(limits=1):
            hinit[...] = L @ np.random.default_rng(100).normal(size=(d, n))
        pout = eng.pinned_empty((2, d + 1, n)); pacc = eng.pinned_empty((2, n), dtype=np.uint8)
        par = amh.MCMCB200(ngpus=k)
        if os.environ.get("AMH_TRACE"):
            print("threads in the process:", [l.split()[1] for l in open("/proc/self/status") if l.startswith("Threads")], "cpus", os.cpu_count(),
                  "affinity", len(os.sched_getaffinity(0)), file=sys.stderr, flush=True)
        for w in range(3):
            amh.sample(model, s, par, 2, n, initial_params=hinit, thinning=spl, chain_type=amh.Chains, seed=99 + w, out=(pout, pacc))
        t0 = time.perf_counter()
        for i in range(reps):
            amh.sample(model, s, par, 2, n, initial_params=hinit, thinning=spl, chain_type=amh.Chains, seed=i, out=(pout, pacc))
        dt = (time.perf_counter() - t0) / reps
        out["e2e"][str(k)] = {"chain_steps_per_s": n * spl / dt, "ms_per_call": dt * 1e3, "chains": n}
        print(f"e2e   ngpus={k}: {n * spl / dt:.4g} chain-steps/s ({dt * 1e3:.2f} ms per sample() call, {n} chains x {spl} steps, host buffers in and out)", flush=True)
        del hinit, pout, pacc
    if ngpus > 1:
        eff = out["e2e"][str(ngpus)]["chain_steps_per_s"] / (ngpus * out["e2e"]["1"]["chain_steps_per_s"])
        out["e2e"]["efficiency"] = eff
        print(f"e2e   weak-scaling efficiency of the in-process job at {ngpus} GPUs: {eff:.3f}", flush=True)


def c5(eng, job, ngpus, out, per_gpu=32768):
    d = 64
    Sg = bench._spd(d, 64, 1e-4, 1.0)
    t = amh.MvNormalTarget(None, Sg)
    n = per_gpu * ngpus
    for tag, S0 in (("from_identity", None), ("from_adapted_S0", (2.38 / np.sqrt(d)) * np.linalg.cholesky(Sg))):
        s = amh.RobustAdaptiveMetropolis() if S0 is None else amh.RobustAdaptiveMetropolis(S=S0)
        run = job.run(job.target_of(t), s.lower(job, d), n, seeds(n, 5), np.zeros((d, n)))
        run.steps(128, warmup=True, steps_per_launch=16)
        rec = {}
        for spl in (1, 16):
            ms = bench._timed(run, 32, warmup=True, spl=spl)
            v = n * 32 / (ms * 1e-3)
            rec[f"warmup_{spl}_per_launch"] = {"chain_steps_per_s": v, "frac_of_hbm_roofline_per_gpu": v / ngpus * 34320 / 1e9 / PEAK}
        st = run.state()
        rec["accept_rate"] = float(st["naccept"].sum() / (n * st["step"]))
        rec["failed_downdates"] = run.ram_failed()[0]
        out["c5"][tag] = rec
        print(f"c5    {n} chains over {ngpus} GPUs, {tag}: " + ", ".join(f"{k} {v['chain_steps_per_s']:.4g} cs/s ({100 * v['frac_of_hbm_roofline_per_gpu']:.1f}%/GPU)"
                                                                        for k, v in rec.items() if isinstance(v, dict)) + f", accept {rec['accept_rate']:.3f}", flush=True)
        run.close()


PEAK = 6537.3
try:
    PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass


def main():
    import torch
    ngpus = int(sys.argv[1]) if len(sys.argv) > 1 else torch.cuda.device_count()
    eng = amh.default_engine(0)
    out = {"ngpus": ngpus, "parity": {}, "broadcast": {}, "e2e": {}, "c5": {}}
    bcast(eng, ngpus, out)
    job = eng.job(ngpus)
    print("job:", ngpus, "devices, broadcast mode", job.broadcast_info(), flush=True)
    parity(eng, job, out)
    c5(eng, job, ngpus, out)
    job.close()
    e2e(eng, ngpus, out)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
