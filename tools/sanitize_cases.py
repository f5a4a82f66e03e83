#!/usr/bin/env python3
"""Small invocations of every hand-synchronised kernel, for `compute-sanitizer --tool racecheck|memcheck|synccheck`
(sanitizers).  Each case is bit-compared with the oracle afterwards, so a sanitizer run is also a
parity run.  Usage: python tools/sanitize_cases.py <case> ; cases:
  c2     K1T16  RWMH d=32 full covariance                       (DMMA tiles in shared memory)
  c3     K2F    stretch, 1-CTA sweep (AMH_STRETCH_CLUSTER=0)    (version flags in shared memory)
  c3cl   K2F    stretch, 2-CTA cluster sweep                    (values pushed into the peer with st.async + mbarriers)
  c3r    K2R    stretch, ensemble resident in a 2-CTA cluster   (in-place records, cluster-barrier windows, st.async pushes)
  c3rs   K2R    same with 100-walker windows, 20 forwarding slots and 4 levels: remote-flag hand-off, overflow bucket
  c4     K3L    MALA logistic d=32, 200 rows                    (TMA producer warp + 10-stage mbarrier ring, wraps)
  c5     K4W    RAM warm-up d=32                                (bulk load / store of the factor, roll-back)
  c5redo K4W    same with the IEEE redo path forced on every step
  c2pad / c2pad60 / c2big  K1T16 zero-padded (d = 13 on 4-warp CTAs, d = 60 as D = 64, d = 29 on the 28-warp CTA)
  c2hast K1     StaticMH with issymmetric = false, exact-dimension kernel with the Hastings term
  c4rw / c4pad / c4rwf  K3L  RWMH (RW variant; c4rwf: full-covariance proposal) / MALA on a logistic regression with 20 features padded to 32
  c4t    K3T    MALA logistic d=128 on the opt-in split-bf16 tcgen05 path (TMA ring, TMEM accumulators, 8 epilogue warps);
                compared with the oracle within the path's stated tolerance instead of bit for bit
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def spd(d, seed, lo, hi):
    rng = np.random.default_rng(seed)
    Q, _ = np.linalg.qr(rng.normal(size=(d, d)))
    lam = np.exp(np.linspace(np.log(lo), np.log(hi), d))
    S = (Q * lam) @ Q.T
    return (S + S.T) / 2


def main():
    case = sys.argv[1]
    if case == "c3":
        os.environ["AMH_STRETCH_RES"] = "0"
        os.environ["AMH_STRETCH_CLUSTER"] = "0"
    if case == "c3cl":
        os.environ["AMH_STRETCH_RES"] = "0"
        os.environ["AMH_STRETCH_CLUSTER"] = "1"
    if case in ("c3r", "c3rs"):
        os.environ["AMH_STRETCH_RES"] = "1"
    if case == "c3rs":
        os.environ["AMH_STRETCH_WIN"] = "100"
        os.environ["AMH_STRETCH_FWD"] = "20"
        os.environ["AMH_STRETCH_LEVELS"] = "4"
    if case == "c5redo":
        os.environ["AMH_RAMW_FORCE_REDO"] = "1"
    if case == "c2big":
        os.environ["AMH_TC_NO_SMALL"] = "1"
    import amh_b200 as amh
    gpu = amh.default_engine(0)
    orc = amh.Engine(lib_path=os.path.join(ROOT, "oracle", "libamh_oracle.so"), prefix="amho_")
    seeds = lambda n, s: np.random.default_rng(s).integers(0, 2 ** 64, size=n, dtype=np.uint64)
    warm, nsteps, spl, init, keys = False, 8, 4, None, ["x", "lp", "accepted", "naccept"]
    if case == "c2":
        d, n = 32, 1024
        Sg = spd(d, 32, 1.0, 100.0)
        t, s, sd = amh.MvNormalTarget(None, Sg), amh.RWMH(amh.MvNormal(np.zeros(d), (2.38 ** 2 / d) * Sg)), seeds(n, 1)
    elif case in ("c2pad", "c2pad60", "c2big"):
        # the zero-padded tensor-core kernels (amh_launch_mh_tcp.cu): d = 13 on 4-warp CTAs (few chains), d = 60 as D = 64 on
        # 16-warp CTAs, and d = 29 with AMH_TC_NO_SMALL: the 28-warp padded kernel; ragged chain counts
        d, n = {"c2pad": (13, 1001), "c2pad60": (60, 777), "c2big": (29, 1501)}[case]
        Sg = spd(d, 32, 1.0, 100.0)
        t, s, sd = amh.MvNormalTarget(np.linspace(-1, 1, d), Sg), amh.RWMH(amh.MvNormal(np.zeros(d), (2.38 ** 2 / d) * Sg)), seeds(n, 1)
    elif case == "c2hast":
        # StaticMH with the reference's default issymmetric = false on an exact-dimension kernel (amh_launch_mh_hast.cu)
        d, n = 9, 700
        Sg = spd(d, 32, 0.5, 2.0)
        t, s, sd = amh.MvNormalTarget(None, Sg), amh.StaticMH(amh.MvNormal(np.linspace(0.1, -0.2, d), 1.5 * Sg)), seeds(n, 1)
    elif case in ("c4rw", "c4pad", "c4rwf"):
        # RWMH (RW variant of K3L) and MALA on a logistic regression with 20 features padded to 32, ragged rows
        d, rows, n = 20, 203, 70
        rng = np.random.default_rng(5)
        X = rng.normal(size=(rows, d)) / np.sqrt(d)
        y = (rng.random(rows) < 0.5).astype(float)
        t = amh.LogisticRegressionTarget(X, y, tau=10.0)
        if case == "c4rw":
            s = amh.RWMH(amh.MvNormal(np.zeros(d), (0.1 ** 2) * amh.I))
        elif case == "c4rwf":                                   # full-covariance proposal: x + L z by DMMA, in-place tile
            s = amh.RWMH(amh.MvNormal(np.zeros(d), (0.1 ** 2 / d) * spd(d, 9, 0.5, 4.0)))
        else:
            s = amh.MALA(lambda g: amh.MvNormal((0.05 / 2) * g, 0.05 * amh.I))
            keys = keys + ["grad"]
        sd, init, nsteps, spl = seeds(n, 3), np.zeros((d, n)), 4, 2
    elif case in ("c3", "c3cl", "c3r", "c3rs"):
        d, nw, ne = 10, 1024, 2
        t = amh.RosenbrockTarget(d)
        s = amh.Ensemble(nw, amh.StretchProposal(amh.MvNormal(np.zeros(d), amh.I)))
        n, sd = nw * ne, seeds(ne, 2)
    elif case == "c4":
        d, rows, n = 32, 200, 64
        rng = np.random.default_rng(5)
        X = rng.normal(size=(rows, d)) / np.sqrt(d)
        y = (rng.random(rows) < 0.5).astype(float)
        t = amh.LogisticRegressionTarget(X, y, tau=10.0)
        s = amh.MALA(lambda g: amh.MvNormal((0.05 / 2) * g, 0.05 * amh.I))
        sd, init, nsteps, spl = seeds(n, 3), np.zeros((d, n)), 4, 2
        keys = keys + ["grad"]
    elif case == "c4t":
        d, rows, n = 128, 300, 200
        rng = np.random.default_rng(5)
        X = rng.normal(size=(rows, d)) / np.sqrt(d)
        y = (rng.random(rows) < 0.5).astype(float)
        t = amh.LogisticRegressionTarget(X, y, tau=10.0)
        s = amh.MALA(lambda g: amh.MvNormal((0.2 / 2) * g, 0.2 * amh.I))
        sd, init, nsteps, spl = seeds(n, 3), 0.05 * rng.normal(size=(d, n)), 1, 1
        keys = ["x", "lp", "grad"]
    elif case in ("c5", "c5redo"):
        d, n = 32, 256
        Sg = spd(d, 64, 1e-2, 1.0)
        t = amh.MvNormalTarget(None, Sg)
        s = amh.RobustAdaptiveMetropolis(S=(2.38 / np.sqrt(d)) * np.linalg.cholesky(Sg))
        sd, init, warm = seeds(n, 4), np.zeros((d, n)), True
        keys = keys + ["S"]
    else:
        raise SystemExit(__doc__)
    res = []
    for eng in (gpu, orc):
        with amh.precision("bf16x2" if (case == "c4t" and eng is gpu) else "fp64"):
            run = eng.run(eng.target(t.kind, d, t.blob()), s.lower(eng, d), n, sd, init)
        run.steps(nsteps, warmup=warm, steps_per_launch=spl)
        res.append(run.state(grad="grad" in keys, S="S" in keys))
        run.close()
    if case == "c4t":
        same = res[0]["accepted"] == res[1]["accepted"]
        assert same.mean() > 0.98 and np.abs(res[0]["lp"][same] - res[1]["lp"][same]).max() < 2e-3
        print(f"{case}: ok, GPU == oracle within the stated tolerance of the bf16x2 path ({n} chains x {nsteps} steps)")
        return
    for k in keys:
        assert np.array_equal(res[0][k], res[1][k]), f"{case}: GPU and oracle differ in {k}"
    print(f"{case}: ok, GPU == oracle bit for bit ({n} chains x {nsteps} steps)")


if __name__ == "__main__":
    main()
