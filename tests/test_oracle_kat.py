"""T0: pin the CPU oracle (oracle/amh_oracle.cpp) before trusting it as the checker.

 * Philox4x32-10 against the Random123 known-answer vectors (kat_vectors of Random123 1.x);
 * the contract's deterministic log/exp/Box-Muller/exponential against mpmath;
 * the reference's OWN statistical known answers, i.e. the assertions of
   /root/reference/test/runtests.jl, test/emcee.jl, test/RobustAdaptiveMetropolis.jl and the RAM doctest.

The reference holds no bit-level golden vectors and Julia is not installed, so bit-level agreement with
a Julia run is UNPINNED (see oracle header)."""
import ctypes as C
import math

import mpmath as mp
import numpy as np
import pytest

from conftest import make_spd


def _seeds(n, s=0):
    return np.random.default_rng(s).integers(0, 2 ** 64, size=n, dtype=np.uint64)


# ------------------------------------------------------------------ RNG + math contract
PHILOX_KAT = [   # ctr[4], key[2] -> out[4]   (Random123 kat_vectors: "philox4x32 10 ...")
    ((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
    ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
    ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
     (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
]


PHILOX7_KAT = [   # Random123 kat_vectors: "philox4x32 7 ..." -- the round count of contract v2's step-noise blocks
    ((0, 0, 0, 0), (0, 0), (0x5f6fb709, 0x0d893f64, 0x4f121f81, 0x4f730a48)),
    ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x5207ddc2, 0x45165e59, 0x4d8ee751, 0x8c52f662)),
    ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
     (0x4dfccaba, 0x190a87f0, 0xc47362ba, 0xb6b5242a)),
]


def test_philox4x32_7_known_answers(oracle):
    out = (C.c_uint32 * 4)()
    for ctr, key, want in PHILOX7_KAT:
        oracle.lib.amho_probe_philox7(*[C.c_uint32(v) for v in ctr], *[C.c_uint32(v) for v in key], out)
        assert tuple(out) == want


def test_contract_v2_step_noise_layout_and_distribution(oracle):
    """contract v2 (include/amh_contract.h): block j of a step -> normals 4j..4j+3 from its four 32-bit words (radius word,
    angle word) x 2, Philox4x32-7; the exponential from the 64-bit word 0 of block ceil(d/4).  Rebuilt here from the raw
    Philox probe with numpy, then checked for normality / independence."""
    import scipy.stats as st
    dp = C.POINTER(C.c_double)
    out = (C.c_uint32 * 4)()
    def block(seed, blk):
        oracle.lib.amho_probe_philox7(C.c_uint32(blk & 0xffffffff), C.c_uint32(blk >> 32), C.c_uint32(0), C.c_uint32(0),
                                      C.c_uint32(seed & 0xffffffff), C.c_uint32(seed >> 32), out)
        return [int(v) for v in out]
    for seed, step, d in ((12345678901234567, 3, 7), (2 ** 63 + 11, 2 ** 33 + 1, 32), (5, 0, 1)):
        z = np.empty(d); e = C.c_double()
        oracle.lib.amho_probe_step_noise_cv(2, C.c_uint64(seed), C.c_uint64(step), d, z.ctypes.data_as(dp), C.byref(e))
        nb = (d + 3) // 4
        want = []
        for j in range(nb):
            v = block(seed, step * (nb + 1) + j)
            for wr, wa in ((v[0], v[1]), (v[2], v[3])):
                u = (wr + 0.5) * 2.0 ** -32
                ang = (np.pi / 2) * ((wa >> 30) + ((wa & 0x3fffffff) * 2.0 ** -30 - 0.5))
                r = np.sqrt(-2.0 * np.log(u))
                want += [r * np.cos(ang), r * np.sin(ang)]
        np.testing.assert_allclose(z, np.array(want)[:d], rtol=1e-13, atol=1e-15)
        v = block(seed, step * (nb + 1) + nb)
        assert e.value == pytest.approx(-np.log((((v[1] << 32 | v[0]) >> 12) + 0.5) * 2.0 ** -52), rel=1e-14)
    zs, es = [], []
    z = np.empty(32); e = C.c_double()
    for seed in _seeds(6000, 77):
        oracle.lib.amho_probe_step_noise_cv(2, C.c_uint64(int(seed)), C.c_uint64(9), 32, z.ctypes.data_as(dp), C.byref(e))
        zs.append(z.copy()); es.append(e.value)
    zs = np.array(zs)
    assert st.kstest(zs.ravel(), "norm").pvalue > 1e-3 and st.kstest(es, "expon").pvalue > 1e-3
    assert abs(zs.var() - 1) < 0.01 and abs(st.kurtosis(zs.ravel())) < 0.05
    cm = np.corrcoef(zs.T)
    assert np.abs(cm - np.eye(32)).max() < 0.06          # no correlation inside or across blocks (6000 samples: sd 0.013)


def test_philox4x32_10_known_answers(oracle):
    out = (C.c_uint32 * 4)()
    for ctr, key, want in PHILOX_KAT:
        oracle.lib.amho_probe_philox(*[C.c_uint32(v) for v in ctr], *[C.c_uint32(v) for v in key], out)
        assert tuple(out) == want


def _probe(oracle, name, x):
    x = np.ascontiguousarray(x, dtype=np.float64)
    y = np.empty_like(x)
    dp = C.POINTER(C.c_double)
    getattr(oracle.lib, name)(x.ctypes.data_as(dp), y.ctypes.data_as(dp), C.c_int64(x.size))
    return y


def _ulp_err(got, want_mp):
    out = []
    for g, w in zip(got, want_mp):
        w64 = float(w)
        if w64 == 0.0 or not math.isfinite(w64):
            out.append(0.0 if g == w64 else np.inf)
            continue
        ulp = math.ulp(w64)
        out.append(abs(float(mp.mpf(g) - w)) / ulp)
    return np.array(out)


def test_log_exp_accuracy_vs_mpmath(oracle):
    mp.mp.prec = 120
    rng = np.random.default_rng(1)
    x = np.concatenate([rng.uniform(0, 1, 2000), np.exp(rng.uniform(-700, 700, 2000)), 1 + rng.uniform(-1e-3, 1e-3, 500),
                        [5e-324, 2.2250738585072014e-308, 1.0, 2.0, 0.5, 1e300]])
    got = _probe(oracle, "amho_probe_log", x)
    err = _ulp_err(got, [mp.log(mp.mpf(float(v))) for v in x])
    assert err.max() < 2.0, err.max()       # table + degree-6 polynomial: < 2 ulp everywhere, < 1 ulp away from x ~ 1
    assert _probe(oracle, "amho_probe_log", [1.0])[0] == 0.0
    sp = _probe(oracle, "amho_probe_log", [0.0, -1.0, np.inf, np.nan])
    assert sp[0] == -np.inf and np.isnan(sp[1]) and sp[2] == np.inf and np.isnan(sp[3])
    x = np.concatenate([rng.uniform(-745, 709, 4000), rng.uniform(-1, 1, 1000), [0.0, -0.0, 709.78, -745.1]])
    got = _probe(oracle, "amho_probe_exp", x)
    err = _ulp_err(got, [mp.exp(mp.mpf(float(v))) for v in x])
    assert err.max() <= 1.0, err.max()      # 1 ulp is reached only on subnormal results
    sp = _probe(oracle, "amho_probe_exp", [800.0, -800.0, np.nan, -np.inf])
    assert sp[0] == np.inf and sp[1] == 0.0 and np.isnan(sp[2]) and sp[3] == 0.0
    x = rng.uniform(-40, 40, 2000)
    got = _probe(oracle, "amho_probe_log1pexp", x)
    err = _ulp_err(got, [mp.log1p(mp.exp(mp.mpf(float(v)))) for v in x])
    assert err.max() < 2.0, err.max()
    got = _probe(oracle, "amho_probe_sigmoid", x)
    err = _ulp_err(got, [1 / (1 + mp.exp(-mp.mpf(float(v)))) for v in x])
    assert err.max() < 3.0, err.max()       # exp (<= 1 ulp) -> 1 + t -> divide -> multiply


def test_uniform_exponential_normal_transforms(oracle):
    mp.mp.prec = 120
    rng = np.random.default_rng(2)
    w = np.concatenate([rng.integers(0, 2 ** 64, size=3000, dtype=np.uint64),
                        np.array([0, 2 ** 64 - 1, 1 << 12, (1 << 12) - 1], dtype=np.uint64)])
    y = np.empty(w.size)
    dp, up = C.POINTER(C.c_double), C.POINTER(C.c_uint64)
    oracle.lib.amho_probe_u01(w.ctypes.data_as(up), y.ctypes.data_as(dp), C.c_int64(w.size))
    want = ((w >> np.uint64(12)).astype(np.float64) + 0.5) * 2.0 ** -52      # exact in fp64
    assert np.array_equal(y, want) and y.min() > 0.0 and y.max() < 1.0
    e = np.empty(w.size)
    oracle.lib.amho_probe_exponential(w.ctypes.data_as(up), e.ctypes.data_as(dp), C.c_int64(w.size))
    err = _ulp_err(e, [-mp.log(mp.mpf(float(v))) for v in want])
    assert err.max() < 2.0 and e.min() > 0.0
    # Box-Muller: z0 = sqrt(-2 ln u) cos(theta), z1 = ... sin(theta), theta = (pi/2)(q + g)
    w0 = rng.integers(0, 2 ** 64, size=3000, dtype=np.uint64)
    w1 = rng.integers(0, 2 ** 64, size=3000, dtype=np.uint64)
    z0 = np.empty(3000); z1 = np.empty(3000)
    oracle.lib.amho_probe_normal_pair(w0.ctypes.data_as(up), w1.ctypes.data_as(up), z0.ctypes.data_as(dp),
                                      z1.ctypes.data_as(dp), C.c_int64(3000))
    for i in range(0, 3000, 7):
        u = (mp.mpf(int(w0[i]) >> 12) + mp.mpf(1) / 2) * mp.mpf(2) ** -52
        q = int(w1[i]) >> 62
        frac = mp.mpf((int(w1[i]) >> 10) & ((1 << 52) - 1)) * mp.mpf(2) ** -52       # in [0,1)
        theta = (mp.pi / 2) * (q + frac - mp.mpf(1) / 2)
        rad = mp.sqrt(-2 * mp.log(u))
        assert abs(float(rad * mp.cos(theta)) - z0[i]) < 1e-14 * max(1.0, abs(z0[i]))
        assert abs(float(rad * mp.sin(theta)) - z1[i]) < 1e-14 * max(1.0, abs(z1[i]))


def test_step_noise_is_standard_normal_and_exponential(oracle):
    d = 32
    z = np.empty(d); e = C.c_double()
    dp = C.POINTER(C.c_double)
    Z, E = [], []
    for seed in range(400):
        for step in (0, 1, 2, 1000003):
            oracle.lib.amho_probe_step_noise(C.c_uint64(seed * 0x9E3779B97F4A7C15 % 2 ** 64), C.c_uint64(step), d,
                                            z.ctypes.data_as(dp), C.byref(e))
            Z.append(z.copy()); E.append(e.value)
    Z = np.array(Z); E = np.array(E)
    n = Z.size
    assert abs(Z.mean()) < 4 / math.sqrt(n) and abs(Z.var() - 1) < 4 * math.sqrt(2 / n)
    assert abs((Z ** 4).mean() - 3) < 0.1
    assert np.abs(np.corrcoef(Z.T) - np.eye(d)).max() < 0.15          # 1600 rows: 4 sigma = 0.1
    assert abs(E.mean() - 1) < 0.1 and E.min() > 0


# ------------------------------------------------- the reference's own statistical known answers
def _run(oracle, target, sampler, n, seeds, init=None):
    th = oracle.target(target.kind, target.dim, target.blob())
    return oracle.run(th, sampler.lower(oracle, target.dim), n, seeds, init)


@pytest.mark.parametrize("ctor", ["static_normals", "static_mvnormal", "static_int", "rw_normals", "rw_mvnormal", "rw_int"])
def test_kat1_iid_normal_model_moments(amh, oracle, ctor):
    """test/runtests.jl:56-94: data = 300 x N(0,1); E[mu] = 0 +- 0.1, E[sigma] = 1 +- 0.1 for the three
    equivalent constructors of StaticMH and RWMH (100 000 draws)."""
    data = np.random.default_rng(1234).normal(0, 1, 300)
    target = amh.IIDNormalTarget(data)
    kind, form = ctor.split("_")
    arg = {"normals": [amh.Normal(0, 1), amh.Normal(0, 1)], "mvnormal": amh.MvNormal(np.zeros(2), amh.I), "int": 2}[form]
    spl = (amh.StaticMH if kind == "static" else amh.RWMH)(arg)
    n = 8                                   # 8 chains x 12 500 saved = 100 000 draws
    run = _run(oracle, target, spl, n, _seeds(n, 7))
    _, _, s = run.sample(12500, discard_initial=500, store=False, store_accepted=False)
    assert abs(s["mean"][0] - 0.0) < 0.1 + abs(data.mean())          # posterior mean of mu = sample mean
    assert abs(s["mean"][1] - 1.0) < 0.1


def test_kat2_symmetric_rw_on_normal_5_07(amh, oracle):
    """test/runtests.jl:253-259: target N(5, 0.7), symmetric random walk, 100 000 draws: mean, std +- 0.05"""
    target = amh.MvNormalTarget(np.array([5.0]), np.array([[0.49]]))
    for P in (amh.SymmetricRandomWalkProposal, amh.RandomWalkProposal):
        spl = amh.MetropolisHastings(P(amh.Normal(0, 1)))
        n = 8
        run = _run(oracle, target, spl, n, _seeds(n, 8))
        _, _, s = run.sample(12500, discard_initial=200, store=False, store_accepted=False)
        assert abs(s["mean"][0] - 5.0) < 0.05 and abs(math.sqrt(s["var"][0]) - 0.7) < 0.05


def test_kat3_mala_issue95_gaussian(amh, oracle):
    """test/runtests.jl:334-365: TheNormalLogDensity(inv(Sigma)), sigma2 = 0.5, 500 000 samples from ones(2):
    mean 0 +- 0.1, cov Sigma atol 0.2"""
    Sigma = np.array([[1.5, 0.35], [0.35, 1.0]])
    target = amh.GaussianPrecisionTarget(np.linalg.inv(Sigma))
    s2 = 0.5
    spl = amh.MALA(lambda g: amh.MvNormal((s2 / 2) * g, s2 * amh.I))
    n = 8
    run = _run(oracle, target, spl, n, _seeds(n, 9), np.ones((2, n)))
    out, _, s = run.sample(62500, store=True, store_accepted=False)
    x = out[:, :2, :].transpose(0, 2, 1).reshape(-1, 2)
    assert np.all(np.abs(x.mean(0)) < 0.1)
    assert np.allclose(np.cov(x.T), Sigma, atol=0.2)
    assert 0.4 < s["accept_rate"] < 0.99


def test_kat3b_mala_basic_iid_model(amh, oracle):
    """test/runtests.jl:288-332: sigma2 = 1e-3, 1 000 samples, discard_initial = 100, initial_params = ones(2)"""
    data = np.random.default_rng(1234).normal(0, 1, 300)
    target = amh.IIDNormalTarget(data)
    s2 = 1e-3
    spl = amh.MALA(lambda g: amh.MvNormal((s2 / 2) * g, s2 * amh.I))
    n = 8
    run = _run(oracle, target, spl, n, _seeds(n, 10), np.ones((2, n)))
    _, _, s = run.sample(1000, discard_initial=100, store=False, store_accepted=False)
    assert abs(s["mean"][0]) < 0.1 + abs(data.mean()) and abs(s["mean"][1] - 1) < 0.1


@pytest.mark.parametrize("log_space", [False, True])
def test_kat4_stretch_move_normal_inverse_gamma(amh, oracle, log_space):
    """test/emcee.jl:3-42 / 44-83: 1 000 walkers x 1 000 iterations; E[s] = 49/24, E[m] = 7/6, +- 0.1"""
    target = amh.NormalInverseGammaToy(log_space=log_space)
    nw = 1000
    rng = np.random.default_rng(100)
    if log_space:
        spl = amh.Ensemble(nw, amh.StretchProposal(amh.MvNormal(np.zeros(2), amh.I)))
        init = None                                                   # init from MvNormal(zeros(2), I) (emcee.jl:62)
    else:
        # the reference's own constructor (emcee.jl:19): initial draw from an array of univariate laws
        spl = amh.Ensemble(nw, amh.StretchProposal([amh.InverseGamma(2, 3), amh.Normal(0, 1)]))
        init = None
    run = _run(oracle, target, spl, nw, _seeds(1, 11), init)
    out, _, s = run.sample(1000, store=True, store_accepted=False)
    sv = np.exp(out[:, 0, :]) if log_space else out[:, 0, :]
    assert abs(sv.mean() - 49 / 24) < 0.1
    assert abs(out[:, 1, :].mean() - 7 / 6) < 0.1


def test_kat5_ram_doctest_covariance(amh, oracle):
    """src/RobustAdaptiveMetropolis.jl:17-70 (doctest): 2-d Gaussian rho = 0.5, 10 000 warm-up + 10 000 samples:
    cov(chain) ~ Sigma rtol 0.2; with eigenvalue bounds [0.1, 2]: norm(cov - Sigma) < 0.2."""
    Sigma = np.array([[1.0, 0.5], [0.5, 1.0]])
    target = amh.MvNormalTarget(None, Sigma)
    for kw in ({}, dict(eigenvalue_lower_bound=0.1, eigenvalue_upper_bound=2.0)):
        spl = amh.RobustAdaptiveMetropolis(**kw)
        n = 8
        run = _run(oracle, target, spl, n, _seeds(n, 12))
        out, _, s = run.sample(10000, discard_initial=10000, num_warmup=10000, store=True, store_accepted=False)
        for c in range(n):
            cov = np.cov(out[:, :2, c].T)
            if kw:
                assert np.linalg.norm(cov - Sigma) < 0.2 * 1.5       # 8 independent chains, each like the doctest's one
            else:
                assert np.allclose(cov, Sigma, rtol=0.3, atol=0.05)
        pooled = np.cov(out[:, :2, :].transpose(0, 2, 1).reshape(-1, 2).T)
        assert np.allclose(pooled, Sigma, rtol=0.2 / 2)
        # mean acceptance probability adapts towards alpha = 0.234 (Vihola 2012)
        st = run.state()
        acc_rate = st["naccept"].sum() / (n * st["step"])
        assert 0.15 < acc_rate < 0.35


@pytest.mark.parametrize("sigma2,which", [(10.0, "upper"), (0.01, "lower")])
def test_kat6_ram_eigenvalue_bounds(amh, oracle, sigma2, which):
    """test/RobustAdaptiveMetropolis.jl:30-71: gamma = 0.51, bounds [0.9, 1.1], 1 000 warm-up steps, every state:
    all eigvals(S) (= diag of the triangular factor) in bounds and the bound is reached (atol 0.05)."""
    target = amh.MvNormalTarget(None, sigma2 * np.eye(2))
    spl = amh.RobustAdaptiveMetropolis(gamma=0.51, eigenvalue_lower_bound=0.9, eigenvalue_upper_bound=1.1)
    n = 4
    run = _run(oracle, target, spl, n, _seeds(n, 13))
    diags = []
    for _ in range(1000):
        run.steps(1, warmup=True)
        S = run.state(S=True)["S"]
        diags.append(S[[0, 2], :].copy())
    diags = np.array(diags)
    assert diags.min() >= 0.9 and diags.max() <= 1.1
    if which == "upper":
        assert np.all(np.abs(diags.max(axis=(0, 1)) - 1.1) < 0.05)
    else:
        assert np.all(np.abs(diags.min(axis=(0, 1)) - 0.9) < 0.05)


def test_kat7_schedule_first_sample_and_ranges(amh, oracle):
    """test/runtests.jl:203-213 (first sample == initial_params) and :125-129 (discard_initial=25, thinning=4)."""
    data = np.random.default_rng(1234).normal(0, 1, 300)
    model = amh.DensityModel(amh.IIDNormalTarget(data))
    val = np.array([0.4, 1.2])
    chain = amh.sample(model, amh.RWMH(2), 10, initial_params=val, engine=oracle)
    assert np.array_equal(chain[0].params, val) and chain[0].accepted is False
    ch = amh.sample(model, amh.RWMH(2), 1000, discard_initial=25, thinning=4, chain_type=amh.Chains, engine=oracle,
                    param_names=["mu", "sigma"])
    assert ch.range() == range(26, 26 + 4 * 1000, 4)
    assert ch.names == ["mu", "sigma", "lp"]
    # the schedule really advances thinning*(N-1)+discard_initial stateful steps
    assert ch.info["summary"] is None or ch.info["summary"]["n_steps"] == 4 * 999 + 25


# ------------------------------------------------- arrays of univariate proposal laws
def test_univariate_family_draws_match_their_laws(amh, oracle):
    """the contract's samplers (Marsaglia-Tsang gamma incl. the shape < 1 boost, inverse gamma, uniform, exponential,
    log-normal) against the analytic moments and scipy's CDFs (Kolmogorov-Smirnov), through the initial draw of a
    StaticProposal over an array of distributions (proposal.jl:26-28, 70-77)"""
    from scipy import stats
    laws = [(amh.InverseGamma(5, 3), stats.invgamma(5, scale=3)), (amh.InverseGamma(2, 3), stats.invgamma(2, scale=3)),
            (amh.Gamma(3, 2), stats.gamma(3, scale=2)), (amh.Gamma(0.5, 2), stats.gamma(0.5, scale=2)),
            (amh.Gamma(1, 1), stats.gamma(1)), (amh.Uniform(-1, 3), stats.uniform(-1, 4)),
            (amh.Exponential(2.5), stats.expon(scale=2.5)), (amh.LogNormal(0.3, 0.5), stats.lognorm(0.5, scale=math.exp(0.3))),
            (amh.Normal(1, 2), stats.norm(1, 2))]
    d, n = len(laws), 100_000
    spl = amh.MetropolisHastings(amh.StaticProposal([l for l, _ in laws]))
    target = amh.MvNormalTarget(None, np.eye(d))
    run = _run(oracle, target, spl, n, _seeds(n, 77))
    x = run.state()["x"]
    for i, (_, ref) in enumerate(laws):
        ks = stats.kstest(x[i], ref.cdf)
        assert ks.statistic < 1.95 / math.sqrt(n), (i, ks)               # alpha ~ 0.001
        m, v = ref.mean(), ref.var()
        if np.isfinite(v):
            assert abs(x[i].mean() - m) < 5 * math.sqrt(v / n), i
    # independent across coordinates (separate sub-streams)
    assert np.abs(np.corrcoef(stats.rankdata(x, axis=1)) - np.eye(d)).max() < 0.02
    # ... and the log-densities the Hastings term uses: a static step from x evaluates logq(x) - logq(c)
    # (checked against scipy through an accept-everything target in test_family_logpdf_matches_scipy)


def test_family_logpdf_matches_scipy(amh, oracle):
    """family_logpdf (the Hastings term of proposal.jl:31-35, 190-192) against scipy.stats: a FLAT target makes
    log alpha = logq(state) - logq(candidate), which is recovered from the accept decisions' threshold by running the
    static sampler over many chains and comparing the acceptance rate with E[min(1, q(x)/q(c))] -- and, exactly, by
    the Python re-evaluation of the ratio for the chains that flipped"""
    from scipy import stats
    laws = [(amh.InverseGamma(2, 3), stats.invgamma(2, scale=3)), (amh.Gamma(0.7, 2), stats.gamma(0.7, scale=2)),
            (amh.Uniform(-1, 3), stats.uniform(-1, 4)), (amh.Exponential(2.5), stats.expon(scale=2.5)),
            (amh.LogNormal(0.3, 0.5), stats.lognorm(0.5, scale=math.exp(0.3))), (amh.Normal(1, 2), stats.norm(1, 2))]
    for law, ref in laws:
        fam, p0, p1, logc = law.component()
        xs = np.concatenate([ref.rvs(size=200, random_state=1), [-0.5, 0.0]])
        got = np.empty(xs.size)
        dp = C.POINTER(C.c_double)
        oracle.lib.amho_probe_family_logpdf(C.c_int32(fam), C.c_double(p0), C.c_double(p1), C.c_double(logc),
                                           xs.ctypes.data_as(dp), got.ctypes.data_as(dp), C.c_int64(xs.size))
        want = ref.logpdf(xs)
        ok = ~np.isposinf(want)            # Gamma(shape < 1) at exactly 0: a pole, measure zero, not drawn
        got, want = got[ok], want[ok]
        fin = np.isfinite(want)
        assert np.array_equal(np.isneginf(got), np.isneginf(want)), type(law).__name__
        assert np.allclose(got[fin], want[fin], rtol=1e-12, atol=1e-12), type(law).__name__


@pytest.mark.parametrize("form", ["array_of_distributions", "array_of_proposals", "namedtuple_mixed_rw"])
def test_kat8_component_proposals_reach_the_known_posterior(amh, oracle, form):
    """README.md:104-112,125-133: `StaticProposal([Normal(0,1), InverseGamma(2,3)])`, an array / NamedTuple of
    proposals, static and random-walk mixed.  Target: the Normal-InverseGamma toy of test/emcee.jl:5-15 whose
    posterior has E[s] = 49/24, E[m] = 7/6 (the reference's own known answer, +- 0.1)"""
    target = amh.NormalInverseGammaToy()
    if form == "array_of_distributions":
        spl = amh.MetropolisHastings(amh.StaticProposal([amh.InverseGamma(2, 3), amh.Normal(0, 1)]))
    elif form == "array_of_proposals":
        spl = amh.MetropolisHastings([amh.StaticProposal(amh.InverseGamma(2, 3)), amh.StaticProposal(amh.Normal(0, 1))])
    else:
        spl = amh.MetropolisHastings(dict(s=amh.StaticProposal(amh.InverseGamma(2, 3)),
                                          m=amh.RandomWalkProposal(amh.Normal(0, 0.8))))
    n = 2000
    run = _run(oracle, target, spl, n, _seeds(n, 5))
    run.steps(300)
    out, _, summ = run.sample(300, store=True, store_accepted=False)
    assert abs(out[:, 0, :].mean() - 49 / 24) < 0.1
    assert abs(out[:, 1, :].mean() - 7 / 6) < 0.1
    assert 0.05 < summ["accept_rate"] < 0.9
