"""T0b: the oracle's ARITHMETIC against independent numerics (numpy / scipy / central differences).

The moment KATs of test_oracle_kat.py cannot see every bug -- MALA with a wrong gradient is still a valid MH chain
with the right stationary law -- so every catalogue target's log-density and gradient, the Hastings terms of every
sampler, the stretch move and RAM's rank-1 Cholesky update / downdate are restated here in a few numpy lines each,
straight from the reference's formulas, and compared with the oracle step by step:

 * log-densities / gradients: scipy.stats and closed forms; gradients also against central differences;
 * one MH / MALA / stretch / RAM step for many chains: candidate, log acceptance ratio and the accept decision
   recomputed from the oracle's own noise probe (amho_probe_step_noise) with scipy log-pdfs
   (mh-core.jl:92-117, proposal.jl:58-64,79-85,190-196, MALA.jl:70-86, emcee.jl:81-93, RAM :123-173);
 * Givens sweeps: S1 S1' - S0 S0' is the rank-1 matrix +/- v v' and S1 equals numpy's Cholesky factor of it
   (LinearAlgebra.lowrankupdate / lowrankdowndate; RAM :153-173);
 * summaries: Welford mean / variance against numpy on a target with |mean| >> std.
"""
import ctypes as C

import numpy as np
import pytest
import scipy.stats as st

from conftest import make_spd

dp = C.POINTER(C.c_double)


def _seeds(n, s=0):
    return np.random.default_rng(s).integers(0, 2 ** 64, size=n, dtype=np.uint64)


def _logp(oracle, th, x):
    oracle.lib.amho_probe_target_logp.restype = C.c_double
    x = np.ascontiguousarray(x, dtype=np.float64)
    return oracle.lib.amho_probe_target_logp(th.h, x.ctypes.data_as(dp))


def _grad(oracle, th, x):
    x = np.ascontiguousarray(x, dtype=np.float64)
    g = np.empty_like(x)
    lp = C.c_double()
    oracle.lib.amho_probe_target_grad(th.h, x.ctypes.data_as(dp), C.byref(lp), g.ctypes.data_as(dp))
    return lp.value, g


def _noise(oracle, seed, step, d, cv=2):
    """the d normals and the exponential of a step under contract version `cv` (the default of new runs is 2)"""
    z = np.empty(d)
    e = C.c_double()
    oracle.lib.amho_probe_step_noise_cv(cv, C.c_uint64(int(seed)), C.c_uint64(step), d, z.ctypes.data_as(dp), C.byref(e))
    return z, e.value


def _logistic_data(rows, d, seed=3):
    rng = np.random.default_rng(seed)
    X = rng.normal(size=(rows, d)) / np.sqrt(d)
    y = (rng.random(rows) < 1 / (1 + np.exp(-X @ rng.normal(size=d)))).astype(float)
    return X, y


# --------------------------------------------------------------------------- targets
def _catalogue(amh):
    """(name, target, independent log-density, independent gradient or None, sample points)"""
    rng = np.random.default_rng(11)
    out = []
    d = 7
    Sg, mu = make_spd(d, 5, 0.5, 30.0), rng.normal(size=d)
    mvn = st.multivariate_normal(mu, Sg)
    out.append(("mvnormal", amh.MvNormalTarget(mu, Sg), mvn.logpdf, lambda x: -np.linalg.solve(Sg, x - mu), rng.normal(size=(6, d)) * 3))
    A = np.linalg.inv(np.array([[1.5, 0.35], [0.35, 1.0]]))          # test/runtests.jl:335-347
    out.append(("gaussprec", amh.GaussianPrecisionTarget(A), lambda x: -0.5 * x @ A @ x, lambda x: -A @ x, rng.normal(size=(6, 2)) * 2))
    dr = 10
    def rosen(x, a=1.0, b=100.0, s=20.0):
        return -np.sum(b * (x[1:] - x[:-1] ** 2) ** 2 + (a - x[:-1]) ** 2) / s
    out.append(("rosenbrock", amh.RosenbrockTarget(dr), rosen, None, rng.normal(size=(6, dr))))
    data = rng.normal(size=300)                                      # test/runtests.jl:23-31
    def iid(th):
        return np.sum(st.norm(th[0], th[1]).logpdf(data)) if th[1] >= 0 else -np.inf
    def iid_grad(th):
        r = data - th[0]
        return np.array([r.sum() / th[1] ** 2, (r ** 2).sum() / th[1] ** 3 - len(data) / th[1]])
    out.append(("iidnormal", amh.IIDNormalTarget(data), iid, iid_grad, np.column_stack([rng.normal(size=6), 0.3 + rng.random(6) * 2])))
    X, y = _logistic_data(157, 12)
    tau = 2.5
    def logistic(b):
        eta = X @ b
        return np.sum(y * eta - np.logaddexp(0.0, eta)) - b @ b / (2 * tau ** 2)
    def logistic_grad(b):
        return X.T @ (y - 1 / (1 + np.exp(-(X @ b)))) - b / tau ** 2
    out.append(("logistic", amh.LogisticRegressionTarget(X, y, tau=tau), logistic, logistic_grad, rng.normal(size=(6, 12)) * 2))
    obs = np.array([1.5, 2.0])                                       # test/emcee.jl:5-15, 46-56
    def nig(th):
        s, m = th
        if not s > 0:
            return -np.inf
        return st.invgamma(2.0, scale=3.0).logpdf(s) + st.norm(0, np.sqrt(s)).logpdf(m) + np.sum(st.norm(m, np.sqrt(s)).logpdf(obs))
    out.append(("nig", amh.NormalInverseGammaToy(obs), nig, None, np.column_stack([0.2 + rng.random(6) * 4, rng.normal(size=6)])))
    out.append(("niglog", amh.NormalInverseGammaToy(obs, log_space=True), lambda th: nig([np.exp(th[0]), th[1]]) + th[0], None,
                np.column_stack([rng.normal(size=6), rng.normal(size=6)])))
    return out


def test_catalogue_log_densities_and_gradients_vs_scipy_and_central_differences(amh, oracle):
    for name, t, f, g, pts in _catalogue(amh):
        th = oracle.target(t.kind, t.dim, t.blob())
        for x in pts:
            want = f(x)
            got = _logp(oracle, th, x)
            assert got == pytest.approx(want, rel=1e-12, abs=1e-11), (name, x)
            if t.has_gradient:
                lp, gr = _grad(oracle, th, x)
                assert lp == pytest.approx(want, rel=1e-12, abs=1e-11), name
                if g is not None:
                    np.testing.assert_allclose(gr, g(x), rtol=1e-11, atol=1e-11, err_msg=name)
                h = 1e-6
                fd = np.array([(f(x + h * e) - f(x - h * e)) / (2 * h) for e in np.eye(t.dim)])
                np.testing.assert_allclose(gr, fd, rtol=2e-6, atol=2e-6 * max(1.0, np.abs(gr).max()), err_msg=name + " (central differences)")
        th.close()
    # support edges (README.md:29-31 uses sigma >= 0; test/emcee.jl:8 uses s > 0)
    for t, x in ((amh.IIDNormalTarget(np.ones(3)), [0.0, -1e-3]), (amh.NormalInverseGammaToy(), [-0.1, 0.0]), (amh.NormalInverseGammaToy(), [0.0, 0.0])):
        th = oracle.target(t.kind, t.dim, t.blob())
        assert _logp(oracle, th, np.array(x)) == -np.inf
        th.close()


# ------------------------------------------------------------------------ one MH step
@pytest.mark.parametrize("kind", ["rw_full", "rw_mean", "static_full", "static_sym"])
def test_mh_step_candidate_logalpha_and_decision_vs_scipy(amh, oracle, kind):
    """mh-core.jl:92-117 with proposal.jl:41-85,190-196, every log-density by scipy"""
    d, n = 5, 400
    Sg = make_spd(d, 21, 0.3, 8.0)
    target = st.multivariate_normal(np.zeros(d), Sg)
    P = make_spd(d, 22, 0.2, 2.0)
    pm = np.linspace(-0.3, 0.4, d) if kind in ("rw_mean", "static_full", "static_sym") else np.zeros(d)
    prop = st.multivariate_normal(pm, P)
    L = np.linalg.cholesky(P)
    dist = amh.MvNormal(pm, P)
    if kind.startswith("rw"):
        spl = amh.RWMH(dist)
    else:
        spl = amh.MetropolisHastings(amh.StaticProposal(dist, issymmetric=True) if kind == "static_sym" else amh.StaticProposal(dist))
    t = amh.MvNormalTarget(None, Sg)
    seeds = _seeds(n, 5)
    x0 = np.random.default_rng(6).normal(size=(d, n))
    run = oracle.run(oracle.target(t.kind, d, t.blob()), spl.lower(oracle, d), n, seeds, x0)
    s0 = run.state()
    run.steps(1)
    s1 = run.state()
    n_acc = 0
    for c in range(n):
        z, e = _noise(oracle, seeds[c], 1, d, run.contract())
        v = pm + L @ z                                             # rand(rng, MvNormal): mu + L z
        x = x0[:, c]
        cand = x + v if kind.startswith("rw") else v
        if kind.startswith("rw"):
            logratio = prop.logpdf(x - cand) - prop.logpdf(cand - x)        # q(p, t, t_cond) = logpdf(p, t - t_cond)
        elif kind == "static_sym":
            logratio = 0.0                                                  # proposal.jl:195-196
        else:
            logratio = prop.logpdf(x) - prop.logpdf(cand)                   # proposal.jl:79-85
        loga = target.logpdf(cand) - s0["lp"][c] + logratio
        assert s0["lp"][c] == pytest.approx(target.logpdf(x), rel=1e-12)
        if abs(loga + e) < 1e-9:
            continue                                               # decision within rounding of the threshold
        acc = -e < loga
        n_acc += acc
        assert bool(s1["accepted"][c]) == acc, (c, loga, e)
        np.testing.assert_allclose(s1["x"][:, c], cand if acc else x, rtol=1e-12, atol=1e-12)
        assert s1["lp"][c] == pytest.approx(target.logpdf(cand) if acc else s0["lp"][c], rel=1e-11)
    assert 0.05 * n < n_acc < 0.95 * n
    run.close()


def test_mala_step_vs_scipy_logpdfs(amh, oracle):
    """MALA.jl:54-93: cand = x + rand(MvNormal(c grad, s2 I)); logratio = q(prop(grad_c), x, cand) - q(prop(grad), cand, x)"""
    d, n, s2 = 12, 300, 0.9
    X, y = _logistic_data(157, d)
    tau = 2.5
    t = amh.LogisticRegressionTarget(X, y, tau=tau)
    def lp_grad(b):
        eta = X @ b
        return np.sum(y * eta - np.logaddexp(0.0, eta)) - b @ b / (2 * tau ** 2), X.T @ (y - 1 / (1 + np.exp(-eta))) - b / tau ** 2
    spl = amh.MALA(lambda g: amh.MvNormal((s2 / 2) * g, s2 * amh.I))
    seeds = _seeds(n, 8)
    x0 = np.random.default_rng(9).normal(size=(d, n)) * 0.3
    run = oracle.run(oracle.target(t.kind, d, t.blob()), spl.lower(oracle, d), n, seeds, x0)
    s0 = run.state(grad=True)
    run.steps(1)
    s1 = run.state(grad=True)
    n_acc = 0
    for c in range(n):
        z, e = _noise(oracle, seeds[c], 1, d, run.contract())
        x = x0[:, c]
        lpx, gx = lp_grad(x)
        np.testing.assert_allclose(s0["grad"][:, c], gx, rtol=1e-10, atol=1e-11)
        cand = x + (s2 / 2) * gx + np.sqrt(s2) * z
        lpc, gc = lp_grad(cand)
        back = st.multivariate_normal((s2 / 2) * gc, s2 * np.eye(d)).logpdf(x - cand)
        fwd = st.multivariate_normal((s2 / 2) * gx, s2 * np.eye(d)).logpdf(cand - x)
        loga = lpc - lpx + back - fwd
        if abs(loga + e) < 1e-8:
            continue
        acc = -e < loga
        n_acc += acc
        assert bool(s1["accepted"][c]) == acc
        np.testing.assert_allclose(s1["x"][:, c], cand if acc else x, rtol=1e-12, atol=1e-13)
        np.testing.assert_allclose(s1["grad"][:, c], gc if acc else gx, rtol=1e-9, atol=1e-10)
    assert 0.2 * n < n_acc < 0.98 * n
    run.close()


def test_stretch_sweep_vs_numpy_loop(amh, oracle):
    """emcee.jl:39-58, 70-102 restated as the reference's own sequential loop; the partner index, z-uniform and
    exponential of every move come from the contract stream (2 blocks per move)."""
    d, nw, ne = 4, 50, 3
    Sg = make_spd(d, 31, 0.5, 4.0)
    target = st.multivariate_normal(np.zeros(d), Sg)
    t = amh.MvNormalTarget(None, Sg)
    a = 2.0
    spl = amh.Ensemble(nw, amh.StretchProposal(amh.MvNormal(np.zeros(d), amh.I), a))
    seeds = _seeds(ne, 12)
    x0 = np.random.default_rng(13).normal(size=(d, nw * ne))
    run = oracle.run(oracle.target(t.kind, d, t.blob()), spl.lower(oracle, d), nw * ne, seeds, x0)
    run.steps(1)
    s1 = run.state()
    ph = (C.c_uint32 * 4)()
    def block(seed, blk):
        oracle.lib.amho_probe_philox(C.c_uint32(blk & 0xffffffff), C.c_uint32(blk >> 32), C.c_uint32(0), C.c_uint32(0),
                                     C.c_uint32(int(seed) & 0xffffffff), C.c_uint32(int(seed) >> 32), ph)
        return [int(v) for v in ph]
    def u01(lo, hi):
        return (((hi << 32 | lo) >> 12) + 0.5) * 2.0 ** -52
    for en in range(ne):
        old = x0[:, en * nw:(en + 1) * nw].T.copy()
        new = old.copy()
        oldlp = target.logpdf(old)
        for i in range(nw):
            b0, b1 = block(seeds[en], (1 * nw + i) * 2), block(seeds[en], (1 * nw + i) * 2 + 1)
            r = ((b0[1] << 32 | b0[0]) * (nw - 1)) >> 64                   # rand(1:(n-1)) - 1
            idx = (i + r + 1) % nw                                          # mod1(i + rand, n), 0-based
            other = new[idx] if idx < i else old[idx]                       # emcee.jl:53
            z = ((a - 1) * u01(b0[2], b0[3]) + 1) ** 2 / a                  # emcee.jl:81
            yv = other + z * (old[i] - other)
            alpha = (d - 1) * np.log(z) + target.logpdf(yv) - oldlp[i]
            e = -np.log(u01(b1[0], b1[1]))
            if abs(alpha + e) < 1e-9:
                new[i] = s1["x"][:, en * nw + i]
                continue
            acc = -e <= alpha                                               # emcee.jl:93
            new[i] = yv if acc else old[i]
            assert bool(s1["accepted"][en * nw + i]) == acc
        np.testing.assert_allclose(s1["x"][:, en * nw:(en + 1) * nw].T, new, rtol=1e-12, atol=1e-13)
    run.close()


# ------------------------------------------------------------------- RAM: Givens sweeps
@pytest.mark.parametrize("d", [2, 9, 33])
def test_ram_adaptation_is_the_rank_one_cholesky_update_or_downdate(amh, oracle, d):
    """RAM :123-173: x_new = S U + x; eta = k^-gamma; S1 S1' = S0 S0' + sign(da) eta |da| (S0 U)(S0 U)' / |U|^2,
    S1 = numpy's Cholesky factor of that matrix (lowrankupdate / lowrankdowndate keep the diagonal positive)."""
    n, alpha, gamma = 64, 0.234, 0.6
    Sg = make_spd(d, 41, 0.05, 3.0)
    target = st.multivariate_normal(np.zeros(d), Sg)
    t = amh.MvNormalTarget(None, Sg)
    S0 = (1.6 / np.sqrt(d)) * np.linalg.cholesky(0.5 * Sg + 0.5 * make_spd(d, 42, 0.05, 3.0))   # acceptance near the target: both signs occur
    spl = amh.RobustAdaptiveMetropolis(alpha=alpha, gamma=gamma, S=S0)
    seeds = _seeds(n, 14)
    x0 = np.random.default_rng(15).normal(size=(d, n)) * 0.5
    run = oracle.run(oracle.target(t.kind, d, t.blob()), spl.lower(oracle, d), n, seeds, x0)
    tril = np.tril_indices(d)
    prev = run.state(S=True)
    ups = downs = 0
    for k in (1, 2, 3):
        run.steps(1, warmup=True)
        cur = run.state(S=True)
        for c in range(n):
            U, e = _noise(oracle, seeds[c], k, d, run.contract())
            S = np.zeros((d, d)); S[tril] = prev["S"][:, c]
            S1 = np.zeros((d, d)); S1[tril] = cur["S"][:, c]
            x = prev["x"][:, c]
            xn = S @ U + x
            la = min(target.logpdf(xn) - prev["lp"][c], 0.0)
            assert cur["logalpha"][c] == pytest.approx(la, rel=1e-10, abs=1e-10)
            assert cur["eta"][c] == pytest.approx(k ** -gamma, rel=1e-14)
            if abs(e + la) > 1e-9:
                acc = e > -la                                       # RAM :148
                assert bool(cur["accepted"][c]) == acc
                np.testing.assert_allclose(cur["x"][:, c], xn if acc else x, rtol=1e-12, atol=1e-13)
            da = np.exp(la) - alpha
            v = np.sqrt(k ** -gamma * abs(da)) * (S @ U) / np.linalg.norm(U)
            M = S @ S.T + np.sign(da) * np.outer(v, v)
            np.testing.assert_allclose(S1 @ S1.T, M, rtol=1e-11, atol=1e-12)
            np.testing.assert_allclose(S1, np.linalg.cholesky(M), rtol=1e-9, atol=1e-11)
            ups += da > 0
            downs += da < 0
        prev = cur
    assert ups > 0 and downs > 0
    assert not cur["failed"].any()
    run.close()


# -------------------------------------------------------------------- summaries (Welford)
def test_welford_summary_matches_numpy_even_when_mean_dwarfs_std(amh, oracle):
    """running mean / M2 per chain, pooled on the host.  Target N(1e8, 1e-4^2): E[x^2] - m^2 would
    lose every digit of the variance (1e16 vs 1e-8); Welford keeps them."""
    d, n, N = 2, 48, 200
    mu = np.array([1e8, -3e7])
    Sg = np.diag([1e-8, 4e-8])
    t = amh.MvNormalTarget(mu, Sg)
    spl = amh.RWMH(amh.MvNormal(np.zeros(d), 2.0 * Sg))
    x0 = mu[:, None] + np.random.default_rng(1).normal(size=(d, n)) * 1e-4
    run = oracle.run(oracle.target(t.kind, d, t.blob()), spl.lower(oracle, d), n, _seeds(n, 2), x0)
    out, acc, summ = run.sample(N, 10, 3, 0, store=True, summary=True, chain_means=True)
    xs = out[:, :d, :]
    sh = xs - mu[None, :, None]                                       # exact shift: the reference numbers lose nothing
    np.testing.assert_allclose(summ["mean"] - mu, sh.mean(axis=(0, 2)), rtol=1e-4, atol=1e-7)       # a few ulp of 1e8
    np.testing.assert_allclose(summ["chain_mean"] - mu[:, None], sh.mean(axis=0), rtol=1e-3, atol=2e-7)
    np.testing.assert_allclose(summ["var"], sh.var(axis=(0, 2)), rtol=1e-4)     # limited by the ulp of the mean (1.5e-8), not by cancellation
    naive = (xs ** 2).mean(axis=(0, 2)) - xs.mean(axis=(0, 2)) ** 2            # what sum / sum-of-squares accumulators give
    assert (np.abs(naive - sh.var(axis=(0, 2))) > 0.5 * sh.var(axis=(0, 2))).all()
    assert (summ["var"] > 1e-9).all() and (summ["var"] < 1e-6).all()
    assert summ["accept_rate"] == pytest.approx(acc[1:].mean(), abs=0.05)
    run.close()


def test_seed_blocks_do_not_depend_on_the_sharding():
    """sampling._draw_seeds: chain c gets draw number c of the caller's generator whichever rank asks (jumpable PCG64)"""
    from amh_b200 import sampling
    full = np.random.default_rng(77).integers(0, 2 ** 64, size=1000, dtype=np.uint64)
    for lo, hi in ((0, 1000), (0, 1), (123, 700), (999, 1000), (500, 500)):
        rng = np.random.default_rng(77)
        np.testing.assert_array_equal(sampling._draw_seeds(rng, 1000, lo, hi), full[lo:hi])
        after = rng.integers(0, 2 ** 64, size=3, dtype=np.uint64)       # the caller's generator has moved past all 1000
        ref = np.random.default_rng(77)
        ref.integers(0, 2 ** 64, size=1000, dtype=np.uint64)
        np.testing.assert_array_equal(after, ref.integers(0, 2 ** 64, size=3, dtype=np.uint64))
    rng = np.random.Generator(np.random.MT19937(5))                      # not jumpable by draws: falls back to the full draw
    ref = np.random.Generator(np.random.MT19937(5)).integers(0, 2 ** 64, size=50, dtype=np.uint64)
    np.testing.assert_array_equal(sampling._draw_seeds(rng, 50, 10, 20), ref[10:20])
