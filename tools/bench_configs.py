#!/usr/bin/env python3
"""Throughput of the BASELINE configs C2..C5 through the C ABI on one GPU (parity-test shapes, not bench lines).
Lengths are shortened; chain-steps/s is per-step and length independent after warm-up.
Usage: python tools/bench_configs.py [c2 c3 c4 c5 ...]"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import amh_b200 as amh   # noqa: E402

PEAK = 6551.7
try:
    PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass


def spd(d, seed, lo, hi):
    rng = np.random.default_rng(seed)
    Q, _ = np.linalg.qr(rng.normal(size=(d, d)))
    lam = np.exp(np.linspace(np.log(lo), np.log(hi), d))
    S = (Q * lam) @ Q.T
    return (S + S.T) / 2


def timed(run, nsteps, warmup=False, spl=0, reps=3, warm=1):
    for _ in range(warm):
        run.steps(nsteps, warmup=warmup, steps_per_launch=spl)
    run.sync()
    run.kernel_time_ms(reset=True)
    for _ in range(reps):
        run.steps(nsteps, warmup=warmup, steps_per_launch=spl)
    run.sync()
    ms, nl = run.kernel_time_ms(reset=True)
    return ms / reps


def report(name, nchain_steps, ms, bytes_per, extra=""):
    v = nchain_steps / (ms * 1e-3)
    gbs = v * bytes_per / 1e9
    print(f"{name:34s} {v:10.4g} chain-steps/s   {gbs:8.1f} GB/s algorithmic = {100 * gbs / PEAK:5.1f}% of {PEAK:.0f}  {extra}", flush=True)


def main():
    which = sys.argv[1:] or ["c2", "c3", "c4", "c5"]
    eng = amh.Engine(lib_path=os.environ["AMH_LIB"]) if os.environ.get("AMH_LIB") else amh.default_engine(0)
    seeds = lambda n, s: np.random.default_rng(s).integers(0, 2 ** 64, size=n, dtype=np.uint64)
    if "c2" in which:
        for d in [int(v) for v in os.environ.get("AMH_BENCH_DIMS", "32,24,16,10,2").split(",")]:
            n = int(os.environ.get("AMH_BENCH_N", "65536"))
            Sigma = spd(d, 32, 1.0, 100.0)
            t = amh.MvNormalTarget(None, Sigma)
            s = amh.RWMH(amh.MvNormal(np.zeros(d), (2.38 ** 2 / d) * Sigma))
            run = eng.run(eng.target(t.kind, d, t.blob()), s.lower(eng, d), n, seeds(n, 1))
            ms = timed(run, 500, spl=500) if os.environ.get("AMH_BENCH_LONG") else timed(run, 200, spl=100)
            st = run.state()
            report(f"C2 RWMH MvNormal d={d} n={n}", n * (500 if os.environ.get("AMH_BENCH_LONG") else 200), ms, 2 * (d + 1) * 8, f"accept={st['naccept'].sum() / (n * st['step']):.3f}")
            run.close()
    if "c2iso" in which:
        # the isotropic / diagonal proposal variant of config 2: K1T16 with c = x + sigma_i z_i instead of the L z mat-vec
        for d, kind in ((32, "scalar"), (32, "diag"), (16, "scalar")):
            n = 65536
            Sigma = spd(d, 32, 1.0, 100.0)
            t = amh.MvNormalTarget(None, Sigma)
            prop = amh.MvNormal(np.zeros(d), (1.2 ** 2 / d) * amh.I) if kind == "scalar" else \
                [amh.Normal(0, 1.2 / np.sqrt(d) * (1 + 0.02 * i)) for i in range(d)]
            s = amh.RWMH(prop)
            run = eng.run(eng.target(t.kind, d, t.blob()), s.lower(eng, d), n, seeds(n, 1))
            ms = timed(run, 500, spl=500)
            st = run.state()
            report(f"C2 RWMH MvNormal d={d} {kind} proposal", n * 500, ms, 2 * (d + 1) * 8, f"accept={st['naccept'].sum() / (n * st['step']):.3f}")
            run.close()
    if "c3" in which:
        d, nw, ne = int(os.environ.get("AMH_C3_D", "10")), int(os.environ.get("AMH_C3_NW", "4096")), int(os.environ.get("AMH_C3_NE", "64"))
        t = amh.MvNormalTarget(None, spd(d, 3, 0.5, 2.0)) if os.environ.get("AMH_C3_TARGET") == "mvn" else amh.RosenbrockTarget(d)
        s = amh.Ensemble(nw, amh.StretchProposal(amh.MvNormal(np.zeros(d), amh.I)))
        run = eng.run(eng.target(t.kind, d, t.blob()), s.lower(eng, d), nw * ne, seeds(ne, 2))
        c3spl = int(os.environ.get("AMH_C3_SPL", "0"))   # 0 = the library's choice
        c3n = max(128, 2 * c3spl)
        ms = timed(run, c3n, spl=c3spl)
        import time as _time                     # the same by the host's clock: the plan kernels run on a second stream
        t0 = _time.perf_counter()
        for _ in range(3):
            run.steps(c3n, steps_per_launch=c3spl)
        run.sync()
        wall_ms = (_time.perf_counter() - t0) * 1e3 / 3
        st = run.state()
        report(f"C3 stretch {os.environ.get('AMH_C3_TARGET', 'Rosenbrock')} d={d} {ne}x{nw} spl={c3spl}", nw * ne * c3n, ms, 2 * (d + 1) * 8, f"accept={st['naccept'].sum() / (nw * ne * st['step']):.3f} events {ms:.3f} ms, host clock {wall_ms:.3f} ms")
        run.close()
    if "c4" in which:
        d, nrows = 128, 10000
        rng = np.random.default_rng(128)
        X = rng.normal(size=(nrows, d)) / np.sqrt(d)
        beta = rng.normal(size=d)
        y = (rng.random(nrows) < 1 / (1 + np.exp(-X @ beta))).astype(float)
        t = amh.LogisticRegressionTarget(X, y, tau=10.0)
        s2 = float(os.environ.get("AMH_C4_S2", "3.3e-2"))       # step size near the MALA optimum for this posterior (sd ~ 0.2 per coordinate)
        s = amh.MALA(lambda g: amh.MvNormal((s2 / 2) * g, s2 * amh.I))
        for n in (16384,):
            run = eng.run(eng.target(t.kind, d, t.blob()), s.lower(eng, d), n, seeds(n, 3), np.zeros((d, n)))
            run.steps(60, steps_per_launch=4)                  # from zeros to the posterior mode region
            st0 = run.state()
            ms = timed(run, 4, spl=2, reps=2, warm=1)
            st = run.state()
            acc = (st['naccept'].sum() - st0['naccept'].sum()) / (n * (st['step'] - st0['step']))
            report(f"C4 MALA logistic d=128 rows=10k n={n}", n * 4, ms, 2 * (2 * d + 1) * 8,
                   f"accept(recent)={acc:.3f}  {5.12e6 * n * 4 / (ms * 1e-3) / 1e12:.2f} TFLOP/s fp64")
            run.close()
    if "c4t" in which:
        # config 4 on the opt-in split-bf16 tcgen05 path (K3T, stated tolerance; csrc/amh_launch_mala_tensor.cu)
        d, nrows = 128, 10000
        rng = np.random.default_rng(128)
        X = rng.normal(size=(nrows, d)) / np.sqrt(d)
        beta = rng.normal(size=d)
        y = (rng.random(nrows) < 1 / (1 + np.exp(-X @ beta))).astype(float)
        t = amh.LogisticRegressionTarget(X, y, tau=10.0)
        s2 = float(os.environ.get("AMH_C4_S2", "3.3e-2"))
        s = amh.MALA(lambda g: amh.MvNormal((s2 / 2) * g, s2 * amh.I))
        for n in (16384,):
            with amh.precision("bf16x2"):
                run = eng.run(eng.target(t.kind, d, t.blob()), s.lower(eng, d), n, seeds(n, 3), np.zeros((d, n)))
            run.steps(60)
            st0 = run.state()
            ms = timed(run, 20, reps=2, warm=1)
            st = run.state()
            acc = (st['naccept'].sum() - st0['naccept'].sum()) / (n * (st['step'] - st0['step']))
            report(f"C4 MALA logistic d=128 rows=10k n={n} bf16x2 tcgen05", n * 20, ms, 2 * (2 * d + 1) * 8,
                   f"accept(recent)={acc:.3f}  {5.12e6 * n * 20 / (ms * 1e-3) / 1e12:.2f} TFLOP/s fp64-equivalent, {3 * 5.12e6 * n * 20 / (ms * 1e-3) / 1e12:.1f} TFLOP/s bf16 issued")
            run.close()
    if "c5" in which:
        d = 64
        Sigma = spd(d, 64, 1e-4, 1.0)
        t = amh.MvNormalTarget(None, Sigma)
        # S0 = I on this target (standard deviations 0.01 .. 1 in 64 dimensions) never gets moving: eta_k = k^-0.6 shrinks
        # log det S by only ~0.3 k^0.4 nats, acceptance is still 0 after 30 000 steps and every step is a downdate.  Start
        # from the usual 2.38/sqrt(d) scaling of the target's factor, so that the timed steps see the stationary mix of
        # rank-1 updates and downdates around the 0.234 target acceptance.
        s = amh.RobustAdaptiveMetropolis(S=(2.38 / np.sqrt(d)) * np.linalg.cholesky(Sigma))
        for n in (32768,):
            run = eng.run(eng.target(t.kind, d, t.blob()), s.lower(eng, d), n, seeds(n, 4), np.zeros((d, n)))
            # a few hundred adaptation steps first
            burn = int(os.environ.get("AMH_C5_BURN", "512"))
            run.steps(burn, warmup=True, steps_per_launch=16)
            st0 = run.state()
            for spl, tag in ((1, "1 step/launch"), (16, "16 steps/launch")):
                ms = timed(run, 32, warmup=True, spl=spl)
                st1 = run.state()
                report(f"C5 RAM warm-up d=64 n={n}", n * 32, ms, 2 * (d + 1) * 8 + d * (d + 1) * 8,
                       f"{tag} accept(recent)={(st1['naccept'].sum() - st0['naccept'].sum()) / (n * max(1, st1['step'] - st0['step'])):.3f}")
                st0 = st1
            for spl, tag in ((1, "1 step/launch"), (16, "16 steps/launch")):
                ms = timed(run, 32, warmup=False, spl=spl)
                st1 = run.state()
                report(f"C5 RAM sampling d=64 n={n}", n * 32, ms, 2 * (d + 1) * 8 + d * (d + 1) // 2 * 8,
                       f"{tag} accept(recent)={(st1['naccept'].sum() - st0['naccept'].sum()) / (n * max(1, st1['step'] - st0['step'])):.3f}")
                st0 = st1
            run.close()


if __name__ == "__main__":
    main()
