"""T1 parity: the CUDA engine against the CPU oracle on the same seeds, through the C ABI.
fp64 paths are required to be BIT-EXACT (states, log-densities, accept flags, counters)."""
import numpy as np
import pytest

from conftest import make_spd

pytestmark = pytest.mark.gpu


def _pair(amh, cuda, oracle, target, sampler, n, seeds, init=None):
    runs = []
    for eng in (cuda, oracle):
        th = eng.target(target.kind, target.dim, target.blob())
        sh = sampler.lower(eng, target.dim)
        runs.append(eng.run(th, sh, n, seeds, init))
    return runs


def _assert_same_state(rg, ro, **kw):
    sg, so = rg.state(**kw), ro.state(**kw)
    for k in ("x", "lp", "accepted", "naccept", "grad", "S"):
        if sg[k] is None:
            continue
        assert np.array_equal(sg[k], so[k], equal_nan=True), f"{k} differs: max abs {np.nanmax(np.abs(sg[k].astype(float) - so[k].astype(float)))}"
    assert sg["step"] == so["step"]


def _seeds(n, s=0):
    return np.random.default_rng(s).integers(0, 2 ** 64, size=n, dtype=np.uint64)


@pytest.mark.parametrize("d,cov", [(1, "scalar"), (2, "scalar"), (2, "diag"), (3, "full"), (8, "full"), (8, "diag"), (13, "full"),
                                   (16, "full"), (24, "full"), (32, "full"), (32, "scalar"), (40, "full"),
                                   (8, "scalar"), (16, "diag"), (24, "scalar"), (32, "diag"),
                                   # dimensions padded to a multiple of 8 on the tensor-core kernels (amh_launch_mh_tcp.cu)
                                   (7, "full"), (9, "full"), (11, "diag"), (14, "full"), (15, "scalar"), (17, "full"), (18, "full"),
                                   (19, "diag"), (21, "full"), (23, "full"), (25, "full"), (27, "scalar"), (28, "full"), (29, "full"),
                                   (31, "full"), (31, "diag"), (33, "full"), (37, "scalar"), (47, "diag"), (48, "full"), (50, "full"),
                                   (56, "full"), (63, "full"), (64, "full"), (64, "scalar"), (65, "full"), (72, "scalar"), (77, "full"), (80, "full"), (90, "full"),
                                   (100, "full"), (112, "scalar"), (127, "full"), (128, "full")])
def test_rwmh_mvnormal_bit_exact(amh, cuda, oracle, d, cov):
    Sigma = make_spd(d, seed=d)
    target = amh.MvNormalTarget(np.linspace(-1, 1, d), Sigma)
    if cov == "scalar":
        prop = amh.MvNormal(np.zeros(d), 0.3 * amh.I)
    elif cov == "diag":
        prop = [amh.Normal(0, 0.5 + 0.1 * i) for i in range(d)]
    else:
        prop = amh.MvNormal(np.zeros(d), (2.38 ** 2 / d) * Sigma)
    spl = amh.RWMH(prop)
    n = 1000
    rg, ro = _pair(amh, cuda, oracle, target, spl, n, _seeds(n, d))
    _assert_same_state(rg, ro)            # first step: draw from the proposal
    for k, spl_ in [(1, 1), (7, 3), (50, 0)]:
        rg.steps(k, steps_per_launch=spl_)
        ro.steps(k)
        _assert_same_state(rg, ro)
    acc = rg.state()["naccept"].sum() / (n * 58)
    assert 0.02 < acc < 0.98


@pytest.mark.parametrize("sym", [False, True])
@pytest.mark.parametrize("cov", ["scalar", "diag", "full"])
def test_static_mh_bit_exact(amh, cuda, oracle, sym, cov):
    d = 3
    Sigma = make_spd(d, seed=5, lo=0.5, hi=2.0)
    target = amh.MvNormalTarget(np.array([0.2, -0.1, 0.3]), Sigma)
    mean = np.array([0.1, 0.0, -0.2])
    if cov == "scalar":
        dist = amh.MvNormal(mean, 2.0 * amh.I)
    elif cov == "diag":
        dist = [amh.Normal(m, 1.5) for m in mean]
    else:
        dist = amh.MvNormal(mean, 1.5 * Sigma)
    P = amh.SymmetricStaticProposal if sym else amh.StaticProposal
    spl = amh.MetropolisHastings(P(dist))
    n = 512
    rg, ro = _pair(amh, cuda, oracle, target, spl, n, _seeds(n, 11))
    _assert_same_state(rg, ro)
    for k in (1, 20):
        rg.steps(k)
        ro.steps(k)
        _assert_same_state(rg, ro)


@pytest.mark.parametrize("kind,d,cov", [("mvnormal", 1, "scalar"), ("mvnormal", 2, "full"), ("mvnormal", 5, "diag"), ("mvnormal", 8, "full"),
                                        ("gaussprec", 9, "full"), ("rosenbrock", 10, "scalar"), ("mvnormal", 14, "full"),
                                        ("mvnormal", 16, "full"), ("gaussprec", 20, "diag"), ("mvnormal", 24, "full"),
                                        ("rosenbrock", 28, "full"), ("mvnormal", 32, "full"), ("mvnormal", 13, "full"),
                                        ("mvnormal", 40, "full")])
def test_static_mh_default_asymmetric_exact_dimension_kernels(amh, cuda, oracle, kind, d, cov):
    """`StaticMH(dist)` as the reference builds it (StaticProposal{false}: the Hastings term logq(state) - logq(cand) is
    evaluated every step, mh-core.jl:119-123) on the exact-dimension kernels of amh_launch_mh_hast.cu; d = 13 and 40 stay
    on the generic kernel.  States, cached logq (through the accept decisions), samples and a state round trip."""
    Sigma = make_spd(d, seed=20 + d, lo=0.5, hi=2.0)
    if kind == "mvnormal":
        target = amh.MvNormalTarget(np.linspace(0.2, -0.1, d), Sigma)
    elif kind == "gaussprec":
        target = amh.GaussianPrecisionTarget(np.linalg.inv(Sigma))
    else:
        target = amh.RosenbrockTarget(d)
    mean = np.linspace(0.1, -0.2, d) if kind != "rosenbrock" else np.ones(d)
    if cov == "scalar":
        dist = amh.MvNormal(mean, (0.3 if kind == "rosenbrock" else 2.0) * amh.I)
    elif cov == "diag":
        dist = [amh.Normal(m, 1.5) for m in mean]
    else:
        dist = amh.MvNormal(mean, (0.05 if kind == "rosenbrock" else 1.5) * Sigma)
    spl = amh.StaticMH(dist)
    n = 600
    rg, ro = _pair(amh, cuda, oracle, target, spl, n, _seeds(n, 500 + d))
    _assert_same_state(rg, ro)
    for k, spl_ in [(1, 1), (9, 4), (40, 0)]:
        rg.steps(k, steps_per_launch=spl_)
        ro.steps(k)
        _assert_same_state(rg, ro)
    og, ag, _ = rg.sample(5, 2, 3)
    oo, ao, _ = ro.sample(5, 2, 3)
    assert np.array_equal(og, oo) and np.array_equal(ag, ao)
    rg.set_state(rg.state()); ro.set_state(ro.state())        # logq(state) is recomputed on restore
    rg.steps(11); ro.steps(11)
    _assert_same_state(rg, ro)


def test_rw_nonzero_mean_hastings_bit_exact(amh, cuda, oracle):
    d = 4
    Sigma = make_spd(d, seed=9, lo=0.5, hi=2.0)
    target = amh.MvNormalTarget(None, Sigma)
    spl = amh.MetropolisHastings(amh.RandomWalkProposal(amh.MvNormal(np.full(d, 0.05), 0.4 * Sigma)))
    n = 256
    rg, ro = _pair(amh, cuda, oracle, target, spl, n, _seeds(n, 12))
    rg.steps(25); ro.steps(25)
    _assert_same_state(rg, ro)


def test_iid_normal_readme_model_bit_exact(amh, cuda, oracle):
    data = np.random.default_rng(1234).normal(0, 1, 30)
    target = amh.IIDNormalTarget(data)
    for spl in (amh.RWMH(2), amh.StaticMH([amh.Normal(0, 1), amh.Normal(0, 1)])):
        n = 300
        init = np.tile(np.array([[0.0], [1.0]]), (1, n))
        rg, ro = _pair(amh, cuda, oracle, target, spl, n, _seeds(n, 3), init)
        _assert_same_state(rg, ro)
        rg.steps(40); ro.steps(40)
        _assert_same_state(rg, ro)


@pytest.mark.parametrize("kind", ["rosenbrock", "gaussprec", "nig", "niglog", "logistic"])
def test_other_targets_bit_exact(amh, cuda, oracle, kind):
    rng = np.random.default_rng(7)
    if kind == "rosenbrock":
        target = amh.RosenbrockTarget(10)
    elif kind == "gaussprec":
        target = amh.GaussianPrecisionTarget(np.linalg.inv(make_spd(6, 3, 0.5, 3.0)))
    elif kind == "nig":
        target = amh.NormalInverseGammaToy()
    elif kind == "niglog":
        target = amh.NormalInverseGammaToy(log_space=True)
    else:
        X = rng.normal(size=(50, 5)) / np.sqrt(5)
        y = (rng.random(50) < 0.5).astype(float)
        target = amh.LogisticRegressionTarget(X, y, tau=3.0)
    d = target.dim
    spl = amh.RWMH(amh.MvNormal(np.zeros(d), 0.05 * amh.I))
    n = 200
    init = rng.normal(size=(d, n)) * 0.3 + (1.0 if kind in ("nig", "rosenbrock") else 0.0)
    rg, ro = _pair(amh, cuda, oracle, target, spl, n, _seeds(n, 21), init)
    _assert_same_state(rg, ro)
    rg.steps(30); ro.steps(30)
    _assert_same_state(rg, ro)


@pytest.mark.parametrize("kind,d,cov", [("gaussprec", 7, "full"), ("rosenbrock", 9, "scalar"), ("gaussprec", 11, "diag"),
                                        ("rosenbrock", 14, "full"), ("gaussprec", 18, "scalar"), ("gaussprec", 28, "full"),
                                        ("rosenbrock", 28, "diag"), ("gaussprec", 13, "full")])
def test_rwmh_other_targets_more_exact_dimensions(amh, cuda, oracle, kind, d, cov):
    """amh_launch_mh_dims.cu: the per-thread MH kernel instantiated for d = 7, 9, 11, 14, 18, 28 (d = 13 stays generic)"""
    Sigma = make_spd(d, seed=d, lo=0.5, hi=4.0)
    target = amh.RosenbrockTarget(d) if kind == "rosenbrock" else amh.GaussianPrecisionTarget(np.linalg.inv(Sigma))
    scale = 0.05 if kind == "rosenbrock" else 1.0
    if cov == "scalar":
        prop = amh.MvNormal(np.zeros(d), (0.3 * scale) ** 2 * amh.I)
    elif cov == "diag":
        prop = [amh.Normal(0, scale * (0.2 + 0.01 * i)) for i in range(d)]
    else:
        prop = amh.MvNormal(np.zeros(d), scale * (2.38 ** 2 / d) * Sigma)
    n = 700
    rg, ro = _pair(amh, cuda, oracle, target, amh.RWMH(prop), n, _seeds(n, 300 + d))
    _assert_same_state(rg, ro)
    for k, spl_ in [(1, 1), (7, 3), (40, 0)]:
        rg.steps(k, steps_per_launch=spl_)
        ro.steps(k)
        _assert_same_state(rg, ro)
    og, ag, _ = rg.sample(4, 1, 3)
    oo, ao, _ = ro.sample(4, 1, 3)
    assert np.array_equal(og, oo) and np.array_equal(ag, ao)


def test_sample_schedule_and_outputs_bit_exact(amh, cuda, oracle):
    d = 5
    Sigma = make_spd(d, seed=2, lo=0.5, hi=4.0)
    target = amh.MvNormalTarget(None, Sigma)
    spl = amh.RWMH(amh.MvNormal(np.zeros(d), 0.5 * Sigma))
    n = 333
    init = np.random.default_rng(0).normal(size=(d, n))
    rg, ro = _pair(amh, cuda, oracle, target, spl, n, _seeds(n, 4), init)
    og, ag, sg = rg.sample(17, discard_initial=0, thinning=1, chain_means=True)
    oo, ao, so = ro.sample(17, discard_initial=0, thinning=1, chain_means=True)
    assert np.array_equal(og[0, :d, :], init)          # first sample is the initial state (test/runtests.jl:203-213)
    assert np.array_equal(og, oo) and np.array_equal(ag, ao)
    og, ag, sg = rg.sample(9, discard_initial=25, thinning=4, chain_means=True)
    oo, ao, so = ro.sample(9, discard_initial=25, thinning=4, chain_means=True)
    assert np.array_equal(og, oo) and np.array_equal(ag, ao)
    for k in ("mean", "var", "chain_mean"):
        assert np.array_equal(sg[k], so[k])
    assert sg["accept_rate"] == so["accept_rate"] and sg["n_steps"] == so["n_steps"] == 16 + 25 + 32


def test_large_config2_shape_bit_exact_and_moments(amh, cuda, oracle):
    """config-2 shape (d=32 full-covariance MvNormal) at 4096 chains against the oracle, then
    65 536 chains on the GPU alone against the analytic moments (T2)."""
    d = 32
    Sigma = make_spd(d, seed=32)
    target = amh.MvNormalTarget(None, Sigma)
    spl = amh.RWMH(amh.MvNormal(np.zeros(d), (2.38 ** 2 / d) * Sigma))
    n = 4096
    rg, ro = _pair(amh, cuda, oracle, target, spl, n, _seeds(n, 32))
    rg.steps(100); ro.steps(100)
    _assert_same_state(rg, ro)
    n = 65536
    th = cuda.target(target.kind, d, target.blob()); sh = spl.lower(cuda, d)
    L = np.linalg.cholesky(Sigma)
    init = L @ np.random.default_rng(1).normal(size=(d, n))        # start in the target law
    run = cuda.run(th, sh, n, _seeds(n, 33), init)
    _, _, s = run.sample(20, discard_initial=200, thinning=50, store=False, store_accepted=False)
    assert 0.15 < s["accept_rate"] < 0.35
    # pooled over 65 536 x 20 draws: relative error of the variances ~ 1e-3
    assert np.allclose(s["var"], np.diag(Sigma), rtol=0.01)
    assert np.all(np.abs(s["mean"]) < 0.01 * np.sqrt(np.diag(Sigma)) * 3)


# ----------------------------------------------------------------------------- MALA (K3)
@pytest.mark.parametrize("kind,d", [("gaussprec", 2), ("mvnormal", 5), ("mvnormal", 16), ("mvnormal", 24),
                                    ("rosenbrock", 10), ("iid", 2), ("logistic", 7),
                                    # amh_launch_mala_dims.cu (exact) and two dimensions that stay generic
                                    ("mvnormal", 6), ("mvnormal", 7), ("gaussprec", 9), ("mvnormal", 12), ("rosenbrock", 14),
                                    ("mvnormal", 20), ("gaussprec", 24), ("mvnormal", 32), ("rosenbrock", 32), ("mvnormal", 11),
                                    ("mvnormal", 40)])
def test_mala_bit_exact(amh, cuda, oracle, kind, d):
    rng = np.random.default_rng(5)
    sigma2 = 0.05
    if kind == "gaussprec":
        # TheNormalLogDensity of test/runtests.jl:335-365
        target = amh.GaussianPrecisionTarget(np.linalg.inv(np.array([[1.5, 0.35], [0.35, 1.0]]) if d == 2 else make_spd(d, seed=d, lo=0.5, hi=4.0)))
        sigma2 = 0.5 if d == 2 else 0.05
    elif kind == "mvnormal":
        target = amh.MvNormalTarget(np.linspace(-0.5, 0.5, d), make_spd(d, seed=d, lo=0.5, hi=4.0))
    elif kind == "rosenbrock":
        target = amh.RosenbrockTarget(d)
        sigma2 = 1e-3
    elif kind == "iid":
        target = amh.IIDNormalTarget(rng.normal(0, 1, 30))
        sigma2 = 1e-3
    else:
        X = rng.normal(size=(40, d)) / np.sqrt(d)
        y = (rng.random(40) < 0.5).astype(float)
        target = amh.LogisticRegressionTarget(X, y, tau=3.0)
    spl = amh.MALA(lambda g: amh.MvNormal((sigma2 / 2) * g, sigma2 * amh.I))
    n = 300
    init = np.ones((d, n)) + 0.05 * rng.normal(size=(d, n))
    rg, ro = _pair(amh, cuda, oracle, target, spl, n, _seeds(n, 40 + d), init)
    _assert_same_state(rg, ro, grad=True)
    for k, spl_ in [(1, 1), (9, 4), (40, 0)]:
        rg.steps(k, steps_per_launch=spl_)
        ro.steps(k)
        _assert_same_state(rg, ro, grad=True)
    acc = rg.state()["naccept"].sum() / (n * 50)
    assert 0.05 < acc <= 1.0


def test_mala_requires_initial_params_and_gradient(amh, cuda):
    target = amh.GaussianPrecisionTarget(np.eye(2))
    spl = amh.MALA(lambda g: amh.MvNormal(0.25 * g, 0.5 * amh.I))
    th = cuda.target(target.kind, 2, target.blob()); sh = spl.lower(cuda, 2)
    with pytest.raises(amh.AMHStateError, match="please specify initial parameters"):      # MALA.jl:37
        cuda.run(th, sh, 4, _seeds(4))
    nig = amh.NormalInverseGammaToy()
    with pytest.raises(amh.AMHArgumentError, match="gradient"):                            # MALA.jl:42-52
        cuda.run(cuda.target(nig.kind, 2, nig.blob()), sh, 4, _seeds(4), np.ones((2, 4)))


# ------------------------------------------------------------------------------ RAM (K4)
@pytest.mark.parametrize("d,bounds", [(2, None), (2, (0.9, 1.1)), (5, None), (16, (0.1, 2.0)), (17, (0.5, 1.5)), (33, None), (64, None), (100, None)])
def test_ram_bit_exact(amh, cuda, oracle, d, bounds):
    Sigma = make_spd(d, seed=60 + d, lo=1e-2, hi=1.0)
    target = amh.MvNormalTarget(None, Sigma)
    kw = {} if bounds is None else dict(eigenvalue_lower_bound=bounds[0], eigenvalue_upper_bound=bounds[1])
    spl = amh.RobustAdaptiveMetropolis(**kw)
    n = 200 if d < 64 else 70
    rg, ro = _pair(amh, cuda, oracle, target, spl, n, _seeds(n, 70 + d))
    _assert_same_state(rg, ro, S=True)                    # x = randn(d), S = I, accepted = true (:175-214)
    assert rg.state()["accepted"].all()
    for k, wu in [(1, True), (30, True), (5, False), (12, True)]:
        rg.steps(k, warmup=wu)
        ro.steps(k, warmup=wu)
        _assert_same_state(rg, ro, S=True)
    S = rg.state(S=True)["S"]
    diag = np.array([S[i * (i + 1) // 2 + i] for i in range(d)])
    assert np.all(diag > 0)
    if bounds is not None:
        assert np.all(diag >= bounds[0]) and np.all(diag <= bounds[1])


def test_ram_user_S_and_init(amh, cuda, oracle):
    d = 3
    target = amh.GaussianPrecisionTarget(np.linalg.inv(make_spd(d, 3, 0.5, 3.0)))
    S0 = np.linalg.cholesky(make_spd(d, 4, 0.2, 1.0))
    spl = amh.RobustAdaptiveMetropolis(alpha=0.3, gamma=0.51, S=S0)
    n = 64
    init = np.random.default_rng(3).normal(size=(d, n))
    rg, ro = _pair(amh, cuda, oracle, target, spl, n, _seeds(n, 5), init)
    rg.steps(25, warmup=True); ro.steps(25, warmup=True)
    _assert_same_state(rg, ro, S=True)
    with pytest.raises(ValueError, match="wrong dimensionality"):          # RAM :202-204
        amh.RobustAdaptiveMetropolis(S=np.eye(4)).lower(cuda, d)


def test_ram_sample_schedule_warmup_bit_exact(amh, cuda, oracle):
    d = 4
    target = amh.MvNormalTarget(None, make_spd(d, 8, 0.1, 2.0))
    spl = amh.RobustAdaptiveMetropolis()
    n = 77
    rg, ro = _pair(amh, cuda, oracle, target, spl, n, _seeds(n, 6))
    og, ag, sg = rg.sample(11, discard_initial=30, thinning=3, num_warmup=37)
    oo, ao, so = ro.sample(11, discard_initial=30, thinning=3, num_warmup=37)
    assert np.array_equal(og, oo) and np.array_equal(ag, ao)
    _assert_same_state(rg, ro, S=True)


# -------------------------------------------------------------------------- stretch (K2)
@pytest.mark.parametrize("kind,d,nw,ne", [("rosenbrock", 10, 64, 3), ("nig", 2, 100, 2), ("niglog", 2, 37, 2),
                                          ("mvnormal", 5, 1024, 2), ("rosenbrock", 10, 4096, 1), ("mvnormal", 20, 50, 2),
                                          # the second translation unit's exact dimensions (amh_launch_stretch_dims.cu) and
                                          # two dimensions that stay on the generic kernels
                                          ("rosenbrock", 6, 200, 2), ("mvnormal", 7, 333, 2), ("rosenbrock", 9, 1024, 1),
                                          ("rosenbrock", 12, 777, 2), ("rosenbrock", 20, 130, 2), ("mvnormal", 24, 96, 2),
                                          ("gaussprec", 12, 150, 2), ("rosenbrock", 11, 64, 2), ("mvnormal", 33, 40, 2),
                                          ("rosenbrock", 14, 300, 2), ("mvnormal", 32, 1100, 1), ("rosenbrock", 32, 96, 2)])
def test_stretch_bit_exact_sequential_sweep(amh, cuda, oracle, kind, d, nw, ne):
    if kind == "rosenbrock":
        target = amh.RosenbrockTarget(d)
    elif kind == "nig":
        target = amh.NormalInverseGammaToy()
    elif kind == "niglog":
        target = amh.NormalInverseGammaToy(log_space=True)
    elif kind == "gaussprec":
        target = amh.GaussianPrecisionTarget(np.linalg.inv(make_spd(d, seed=d, lo=0.5, hi=4.0)))
    else:
        target = amh.MvNormalTarget(None, make_spd(d, seed=d, lo=0.5, hi=4.0))
    spl = amh.Ensemble(nw, amh.StretchProposal(amh.MvNormal(np.zeros(d), amh.I)))
    n = nw * ne
    init = None
    if kind == "nig":
        init = np.abs(np.random.default_rng(2).normal(size=(d, n))) + 0.5
    rg, ro = _pair(amh, cuda, oracle, target, spl, n, _seeds(ne, 90 + d), init)
    _assert_same_state(rg, ro)
    for k, spl_ in [(1, 1), (3, 2), (8, 0)]:
        rg.steps(k, steps_per_launch=spl_)
        ro.steps(k)
        _assert_same_state(rg, ro)
    acc = rg.state()["naccept"].sum() / (n * 12)
    assert 0.05 < acc < 0.99


def test_stretch_sample_bit_exact(amh, cuda, oracle):
    d, nw, ne = 2, 50, 2
    target = amh.NormalInverseGammaToy(log_space=True)
    spl = amh.Ensemble(nw, amh.StretchProposal(amh.MvNormal(np.zeros(d), amh.I)))
    rg, ro = _pair(amh, cuda, oracle, target, spl, nw * ne, _seeds(ne, 8))
    og, ag, sg = rg.sample(13, discard_initial=5, thinning=2)
    oo, ao, so = ro.sample(13, discard_initial=5, thinning=2)
    assert np.array_equal(og, oo) and np.array_equal(ag, ao)
    assert np.array_equal(sg["mean"], so["mean"])


@pytest.mark.parametrize("d,zero_mean", [(7, True), (13, False), (22, True), (30, False), (41, True), (60, False), (70, True),
                                         (101, False)])
def test_padded_tensor_core_path_static_symmetric_sample_and_resume(amh, cuda, oracle, d, zero_mean):
    """the padded K1T16 kernels (amh_launch_mh_tcp.cu): symmetric StaticProposal and RWMH through the sample schedule (the
    save epilogue must skip the padding rows), summaries, zero / non-zero target mean, 1 001 chains (a ragged last warp),
    and a state round trip in the middle (the padding rows of the device state must stay zero)"""
    Sigma = make_spd(d, seed=70 + d, lo=0.5, hi=4.0)
    target = amh.MvNormalTarget(None if zero_mean else np.linspace(0.5, -0.5, d), Sigma)
    n = 1001
    for spl in (amh.MetropolisHastings(amh.SymmetricStaticProposal(amh.MvNormal(np.zeros(d), 1.2 * Sigma))),
                amh.RWMH(amh.MvNormal(np.zeros(d), (2.38 ** 2 / d) * Sigma))):
        rg, ro = _pair(amh, cuda, oracle, target, spl, n, _seeds(n, 78 + d))
        og, ag, sg = rg.sample(9, discard_initial=3, thinning=5, chain_means=True)
        oo, ao, so = ro.sample(9, discard_initial=3, thinning=5, chain_means=True)
        assert np.array_equal(og, oo) and np.array_equal(ag, ao)
        assert np.array_equal(sg["mean"], so["mean"]) and np.array_equal(sg["var"], so["var"])
        _assert_same_state(rg, ro)
        st = rg.state()
        rg.set_state(st); ro.set_state(ro.state())
        rg.steps(33, steps_per_launch=0); ro.steps(33)
        _assert_same_state(rg, ro)


@pytest.mark.parametrize("d,cov", [(8, "full"), (13, "full"), (16, "diag"), (24, "full"), (29, "full"), (32, "full"), (32, "scalar")])
def test_tensor_core_kernels_one_cta_per_sm_shape_with_few_chains(amh, cuda, oracle, monkeypatch, d, cov):
    """runs with few chains take the 4-warp CTA shape (amh_launch_mh_tcp.cu, kSmallRunChains); AMH_TC_NO_SMALL keeps them on
    the 28-warp kernels (exact and padded), which large runs use -- both must reproduce the oracle"""
    monkeypatch.setenv("AMH_TC_NO_SMALL", "1")
    Sigma = make_spd(d, seed=d)
    target = amh.MvNormalTarget(np.linspace(-1, 1, d), Sigma)
    if cov == "scalar":
        prop = amh.MvNormal(np.zeros(d), 0.3 * amh.I)
    elif cov == "diag":
        prop = [amh.Normal(0, 0.2 + 0.02 * i) for i in range(d)]
    else:
        prop = amh.MvNormal(np.zeros(d), (2.38 ** 2 / d) * Sigma)
    n = 1500
    rg, ro = _pair(amh, cuda, oracle, target, amh.RWMH(prop), n, _seeds(n, 900 + d))
    _assert_same_state(rg, ro)
    for k, spl_ in [(1, 1), (7, 3), (30, 0)]:
        rg.steps(k, steps_per_launch=spl_)
        ro.steps(k)
        _assert_same_state(rg, ro)
    og, ag, _ = rg.sample(4, 1, 3)
    oo, ao, _ = ro.sample(4, 1, 3)
    assert np.array_equal(og, oo) and np.array_equal(ag, ao)


@pytest.mark.parametrize("d", [37, 48, 55, 64])
@pytest.mark.parametrize("small", [True, False])
def test_padded_tensor_core_kernels_above_32_both_cta_shapes(amh, cuda, oracle, monkeypatch, d, small):
    """d = 33 ... 64: few chains take 8-warp CTAs, AMH_TC_NO_SMALL keeps the large (24 / 20 / 16-warp) CTAs -- both bit-exact"""
    if not small:
        monkeypatch.setenv("AMH_TC_NO_SMALL", "1")
    Sigma = make_spd(d, seed=d)
    target = amh.MvNormalTarget(np.linspace(-1, 1, d), Sigma)
    n = 900
    rg, ro = _pair(amh, cuda, oracle, target, amh.RWMH(amh.MvNormal(np.zeros(d), (2.38 ** 2 / d) * Sigma)), n, _seeds(n, 1200 + d))
    for k, spl_ in [(1, 1), (7, 3), (20, 0)]:
        rg.steps(k, steps_per_launch=spl_)
        ro.steps(k)
        _assert_same_state(rg, ro)
    og, ag, _ = rg.sample(3, 1, 2)
    oo, ao, _ = ro.sample(3, 1, 2)
    assert np.array_equal(og, oo) and np.array_equal(ag, ao)


def test_tensor_core_path_static_symmetric_and_sample(amh, cuda, oracle):
    """K1T (DMMA mat-vecs) with a symmetric StaticProposal and through the sample schedule / save epilogue"""
    d = 16
    Sigma = make_spd(d, seed=77, lo=0.5, hi=4.0)
    target = amh.MvNormalTarget(np.linspace(0.5, -0.5, d), Sigma)
    spl = amh.MetropolisHastings(amh.SymmetricStaticProposal(amh.MvNormal(np.zeros(d), 1.2 * Sigma)))
    n = 100
    rg, ro = _pair(amh, cuda, oracle, target, spl, n, _seeds(n, 78))
    og, ag, sg = rg.sample(12, discard_initial=3, thinning=2, chain_means=True)
    oo, ao, so = ro.sample(12, discard_initial=3, thinning=2, chain_means=True)
    assert np.array_equal(og, oo) and np.array_equal(ag, ao)
    for k in ("mean", "var", "chain_mean"):
        assert np.array_equal(sg[k], so[k])
    _assert_same_state(rg, ro)


def test_ram_warp_kernel_user_S_fused_steps_and_sample(amh, cuda, oracle):
    """K4W (warp per chain, factor in shared memory): user-supplied S, several fused steps per launch (roll-back and
    write-through of the factor inside one launch), and the sample schedule with warm-up"""
    d = 20
    Sigma = make_spd(d, seed=91, lo=1e-2, hi=1.0)
    target = amh.MvNormalTarget(np.linspace(-0.2, 0.2, d), Sigma)
    S0 = np.linalg.cholesky(make_spd(d, 92, 0.05, 0.5))
    spl = amh.RobustAdaptiveMetropolis(gamma=0.55, S=S0, eigenvalue_lower_bound=0.05, eigenvalue_upper_bound=0.9)
    n = 50
    init = np.random.default_rng(4).normal(size=(d, n)) * 0.1
    rg, ro = _pair(amh, cuda, oracle, target, spl, n, _seeds(n, 93), init)
    _assert_same_state(rg, ro, S=True)
    for k, wu, spl_ in [(7, True, 7), (5, False, 2), (24, True, 5)]:
        rg.steps(k, warmup=wu, steps_per_launch=spl_)
        ro.steps(k, warmup=wu)
        _assert_same_state(rg, ro, S=True)
    og, ag, sg = rg.sample(9, discard_initial=10, thinning=2, num_warmup=50)
    oo, ao, so = ro.sample(9, discard_initial=10, thinning=2, num_warmup=50)
    assert np.array_equal(og, oo) and np.array_equal(ag, ao) and np.array_equal(sg["mean"], so["mean"])
    _assert_same_state(rg, ro, S=True)


@pytest.mark.parametrize("d,rows,n", [(32, 100, 50), (64, 64, 24), (128, 203, 70), (5, 80, 30), (20, 333, 41), (48, 100, 22), (90, 130, 13), (127, 99, 9)])
def test_mala_logistic_tiled_tensor_core_kernel_bit_exact(amh, cuda, oracle, d, rows, n):
    """K3L: the many-row logistic target as two chained DMMA GEMMs (TMA-staged design matrix), against the oracle's
    scalar row loop: candidate, gradient, log-density and accept decisions bit-for-bit; rows % 8 != 0 and
    chains % 8 != 0 exercise the padding paths"""
    rng = np.random.default_rng(d)
    X = rng.normal(size=(rows, d)) / np.sqrt(d)
    beta = rng.normal(size=d)
    y = (rng.random(rows) < 1 / (1 + np.exp(-X @ beta))).astype(float)
    target = amh.LogisticRegressionTarget(X, y, tau=5.0)
    s2 = 0.02
    spl = amh.MALA(lambda g: amh.MvNormal((s2 / 2) * g, s2 * amh.I))
    init = 0.1 * rng.normal(size=(d, n))
    rg, ro = _pair(amh, cuda, oracle, target, spl, n, _seeds(n, 300 + d), init)
    _assert_same_state(rg, ro, grad=True)
    for k, spl_ in [(1, 1), (5, 2), (12, 0)]:
        rg.steps(k, steps_per_launch=spl_)
        ro.steps(k)
        _assert_same_state(rg, ro, grad=True)
    acc = rg.state()["naccept"].sum() / (n * 18)
    assert 0.1 < acc <= 1.0
    og, ag, sg = rg.sample(5, discard_initial=2, thinning=2)
    oo, ao, so = ro.sample(5, discard_initial=2, thinning=2)
    assert np.array_equal(og, oo) and np.array_equal(ag, ao)


@pytest.mark.parametrize("d,rows,n,cov", [(32, 100, 50, "scalar"), (64, 64, 24, "diag"), (128, 203, 70, "scalar"), (128, 1000, 33, "diag"),
                                          (32, 77, 40, "int"),
                                          # features padded to 32 / 64 / 128 with zero columns of the design matrix
                                          (3, 90, 30, "scalar"), (7, 64, 25, "diag"), (20, 333, 41, "scalar"), (33, 100, 20, "diag"),
                                          (50, 150, 26, "scalar"), (100, 120, 17, "diag"), (127, 99, 9, "scalar"),
                                          # full-covariance proposal: candidate = x + L z as a DMMA mat-vec inside the kernel
                                          (32, 100, 50, "full"), (20, 203, 41, "full"), (64, 80, 13, "full"), (77, 90, 9, "full"),
                                          (128, 150, 21, "full")])
def test_rwmh_logistic_tiled_tensor_core_kernel_bit_exact(amh, cuda, oracle, d, rows, n, cov):
    """the RW variant of K3L (amh_launch_mala_logistic.cu, RW = true): RWMH with an isotropic / diagonal proposal on the
    many-row logistic target -- GEMM1 and the log-likelihood terms on the FP64 tensor cores, no gradient -- against the
    oracle's scalar row loop, incl. `RWMH(d::Int)`, ragged rows / chains, the sample schedule and a state round trip"""
    rng = np.random.default_rng(1000 + d)
    X = rng.normal(size=(rows, d)) / np.sqrt(d)
    beta = rng.normal(size=d)
    y = (rng.random(rows) < 1 / (1 + np.exp(-X @ beta))).astype(float)
    target = amh.LogisticRegressionTarget(X, y, tau=5.0)
    if cov == "scalar":
        spl = amh.RWMH(amh.MvNormal(np.zeros(d), (0.08 ** 2) * amh.I))
    elif cov == "diag":
        spl = amh.RWMH([amh.Normal(0, 0.05 + 0.0005 * i) for i in range(d)])
    elif cov == "full":
        spl = amh.RWMH(amh.MvNormal(np.zeros(d), (0.1 ** 2 / d) * make_spd(d, seed=3 * d, lo=0.5, hi=4.0)))
    else:
        spl = amh.RWMH(d)                                       # MvNormal(Zeros(d), I): mh-core.jl:51
    init = 0.1 * rng.normal(size=(d, n))
    rg, ro = _pair(amh, cuda, oracle, target, spl, n, _seeds(n, 700 + d), init)
    _assert_same_state(rg, ro)
    for k, spl_ in [(1, 1), (5, 2), (12, 0)]:
        rg.steps(k, steps_per_launch=spl_)
        ro.steps(k)
        _assert_same_state(rg, ro)
    og, ag, sg = rg.sample(5, discard_initial=2, thinning=2)
    oo, ao, so = ro.sample(5, discard_initial=2, thinning=2)
    assert np.array_equal(og, oo) and np.array_equal(ag, ao)
    rg.set_state(rg.state()); ro.set_state(ro.state())
    rg.steps(7); ro.steps(7)
    _assert_same_state(rg, ro)


@pytest.mark.parametrize("warps", ["3", "8", "14"])
@pytest.mark.parametrize("sampler", ["mala", "rwmh"])
def test_logistic_tiled_kernels_multi_warp_ctas(amh, cuda, oracle, monkeypatch, warps, sampler):
    """runs with few chains get one consumer warp per CTA (the chain groups are spread over the SMs); AMH_K3L_WARPS forces the
    CTA shapes large runs use (several consumer warps sharing one TMA ring of X) -- same bits"""
    monkeypatch.setenv("AMH_K3L_WARPS", warps)
    d, rows, n = 32, 150, 8 * 14 * 2 + 5
    rng = np.random.default_rng(77)
    X = rng.normal(size=(rows, d)) / np.sqrt(d)
    y = (rng.random(rows) < 1 / (1 + np.exp(-X @ rng.normal(size=d)))).astype(float)
    target = amh.LogisticRegressionTarget(X, y, tau=5.0)
    s2 = 0.02
    spl = amh.MALA(lambda g: amh.MvNormal((s2 / 2) * g, s2 * amh.I)) if sampler == "mala" else amh.RWMH(amh.MvNormal(np.zeros(d), (0.08 ** 2) * amh.I))
    rg, ro = _pair(amh, cuda, oracle, target, spl, n, _seeds(n, 333), 0.1 * rng.normal(size=(d, n)))
    for k, spl_ in [(1, 1), (6, 2)]:
        rg.steps(k, steps_per_launch=spl_)
        ro.steps(k)
        _assert_same_state(rg, ro, grad=sampler == "mala")


def test_sample_pipelined_copy_many_chunks_and_pinned_buffers(amh, cuda, oracle):
    """amh_run_sample drains the device sample ring through the copy stream in chunks (two buffers): force many
    chunks (N large, small slabs) and use caller-owned pinned buffers; results must equal the oracle's"""
    d = 3
    Sigma = make_spd(d, seed=12, lo=0.5, hi=2.0)
    target = amh.MvNormalTarget(None, Sigma)
    spl = amh.RWMH(amh.MvNormal(np.zeros(d), 0.5 * Sigma))
    n = 40
    rg, ro = _pair(amh, cuda, oracle, target, spl, n, _seeds(n, 13))
    N = 257
    pout = cuda.pinned_empty((N, d + 1, n)); pacc = cuda.pinned_empty((N, n), dtype=np.uint8)
    og, ag, _ = rg.sample(N, discard_initial=3, thinning=1, out=pout, acc=pacc)
    oo, ao, _ = ro.sample(N, discard_initial=3, thinning=1)
    assert og is pout and np.array_equal(og, oo) and np.array_equal(ag, ao)
    # strided initial parameters (a column block of a larger matrix) go through init_ld without repacking
    big = np.random.default_rng(5).normal(size=(d, 100))
    r1 = cuda.run(cuda.target(target.kind, d, target.blob()), spl.lower(cuda, d), n, _seeds(n, 14), big[:, 30:70])
    r2 = oracle.run(oracle.target(target.kind, d, target.blob()), spl.lower(oracle, d), n, _seeds(n, 14), np.ascontiguousarray(big[:, 30:70]))
    r1.steps(5); r2.steps(5)
    _assert_same_state(r1, r2)
    assert np.array_equal(r1.state()["x"].shape, (d, n))


def _resume_cases(amh):
    """(name, target, sampler, n, nseeds, init) -- one case per state layout amh_run_set_state must restore"""
    rng = np.random.default_rng(5)
    S5 = make_spd(5, 2, 0.5, 4.0)
    S32 = make_spd(32, 32)
    S20 = make_spd(20, 6, 0.01, 1.0)
    cases = [
        ("rw_k1", amh.MvNormalTarget(None, S5), amh.RWMH(amh.MvNormal(np.zeros(5), 0.5 * S5)), 130, 130, None),
        ("rw_k1t16", amh.MvNormalTarget(None, S32), amh.RWMH(amh.MvNormal(np.zeros(32), (2.38 ** 2 / 32) * S32)), 200, 200, None),
        ("static_lq", amh.MvNormalTarget(None, S5), amh.MetropolisHastings(amh.StaticProposal(amh.MvNormal(np.full(5, 0.1), 1.5 * S5))), 130, 130, None),
        ("mala", amh.GaussianPrecisionTarget(np.linalg.inv(S5)), amh.MALA(lambda g: amh.MvNormal(0.1 * g, 0.2 * amh.I)), 100, 100,
         rng.normal(size=(5, 100))),
        ("ram_k4", amh.MvNormalTarget(None, S5), amh.RobustAdaptiveMetropolis(), 90, 90, None),
        ("ram_k4w", amh.MvNormalTarget(None, S20), amh.RobustAdaptiveMetropolis(eigenvalue_lower_bound=0.05, eigenvalue_upper_bound=3.0), 70, 70, None),
        ("stretch", amh.RosenbrockTarget(4), amh.Ensemble(50, amh.StretchProposal(amh.MvNormal(np.zeros(4), amh.I))), 150, 3, None),
    ]
    return cases


@pytest.mark.parametrize("case", range(7))
def test_set_state_resume_bit_exact(amh, cuda, oracle, case):
    """amh_run_set_state: a fresh run that is handed the state of another one continues it bit for bit
    (initial_state resume; state structs src/AdvancedMH.jl:61-65, MALA.jl:14-19, RAM :99-114), on the GPU and in the oracle"""
    name, target, spl, n, ns, init = _resume_cases(amh)[case]
    seeds = _seeds(ns, 40 + case)
    is_ram = isinstance(spl, amh.RobustAdaptiveMetropolis)
    kw = dict(grad=isinstance(spl, amh.MALA), S=is_ram)
    rg, ro = _pair(amh, cuda, oracle, target, spl, n, seeds, init)
    for r in (rg, ro):
        r.steps(13, warmup=is_ram)              # odd number of sweeps: the stretch double buffer is swapped
    _assert_same_state(rg, ro, **kw)
    saved = rg.state(**kw)
    # fresh handles started somewhere else entirely
    other = np.random.default_rng(1).normal(size=(target.dim, n))
    rg2, ro2 = _pair(amh, cuda, oracle, target, spl, n, seeds, other)
    for r in (rg2, ro2):
        r.set_state(saved)
        assert r.state()["step"] == 13
    _assert_same_state(rg2, rg, **kw)
    for r in (rg, ro, rg2, ro2):
        r.steps(6, warmup=is_ram)
        r.steps(5, warmup=False)
    _assert_same_state(rg2, rg, **kw)
    _assert_same_state(rg2, ro2, **kw)
    _assert_same_state(rg, ro, **kw)
    if is_ram:
        for a, b in zip(rg2.ram_adapt(), ro.ram_adapt()):
            assert np.array_equal(a, b)
    with pytest.raises(ValueError):
        rg.set_state(dict(x=np.zeros((target.dim, n + 1))))
    if not is_ram:
        with pytest.raises(ValueError):
            rg.ram_adapt()
        with pytest.raises(ValueError):
            rg.set_state(dict(S=np.zeros((target.dim * (target.dim + 1) // 2, n))))


def test_sample_initial_state_continues_run(amh, cuda):
    """sample(...; initial_state=) after sample(...; save_state=true) == one long sample call"""
    S = make_spd(6, 8, 0.5, 3.0)
    target = amh.MvNormalTarget(None, S)
    for spl, wu in ((amh.RWMH(amh.MvNormal(np.zeros(6), 0.4 * S)), 0), (amh.RobustAdaptiveMetropolis(), 17)):
        full = amh.sample(np.random.default_rng(9), target, spl, amh.MCMCThreads(), 30, 40, chain_type=amh.Chains, engine=cuda,
                          num_warmup=wu, discard_initial=0)
        a = amh.sample(np.random.default_rng(9), target, spl, amh.MCMCThreads(), 10, 40, chain_type=amh.Chains, engine=cuda,
                       num_warmup=wu, discard_initial=0, save_state=True)
        b = amh.sample(np.random.default_rng(123), target, spl, amh.MCMCThreads(), 20, 40, chain_type=amh.Chains, engine=cuda,
                       num_warmup=wu, discard_initial=0, initial_state=a.info["state"])
        assert np.array_equal(np.concatenate([a.value, b.value]), full.value)
        assert np.array_equal(np.concatenate([a.accepted, b.accepted]), full.accepted)


def _all_families(amh, d):
    fams = [amh.Normal(0.1, 0.7), amh.InverseGamma(2, 3), amh.Gamma(0.6, 1.5), amh.Gamma(4, 0.5), amh.Uniform(-2, 2),
            amh.Exponential(0.8), amh.LogNormal(-0.2, 0.4)]
    return [fams[i % len(fams)] for i in range(d)]


@pytest.mark.parametrize("form", ["static_array", "static_array_symmetric", "rw_array", "rw_array_symmetric", "mixed",
                                  "mixed_d40", "univariate_invgamma"])
def test_component_proposals_bit_exact(amh, cuda, oracle, form):
    """K1C: arrays of univariate laws / arrays of proposals (proposal.jl:26-35, 132-150, 236-240; README.md:104-133) --
    initial draw, candidate, Hastings term (incl. -Inf / NaN outside a law's support), accept/reject: bit-exact"""
    init = None
    n = 700
    if form.startswith("static_array"):
        target = amh.NormalInverseGammaToy()
        P = amh.SymmetricStaticProposal if form.endswith("symmetric") else amh.StaticProposal
        spl = amh.MetropolisHastings(P([amh.InverseGamma(2, 3), amh.Normal(0, 1)]))
    elif form.startswith("rw_array"):
        # positive-only increment laws: x - c is outside the support -> logpdf = -Inf -> NaN / -Inf log-ratios
        target = amh.MvNormalTarget(None, make_spd(7, 3, 0.5, 4.0))
        P = amh.SymmetricRandomWalkProposal if form.endswith("symmetric") else amh.RandomWalkProposal
        laws = _all_families(amh, 7)
        if not form.endswith("symmetric"):
            laws = [amh.Normal(0.05, 0.3), amh.Uniform(-0.5, 0.4), amh.Normal(-0.02, 0.2), amh.Uniform(-0.3, 0.3),
                    amh.Normal(0, 0.3), amh.Uniform(-0.2, 0.25), amh.Normal(0.0, 0.25)]
        spl = amh.MetropolisHastings(P(laws))
    elif form == "mixed":
        target = amh.NormalInverseGammaToy()
        spl = amh.MetropolisHastings([amh.StaticProposal(amh.InverseGamma(2, 3)), amh.RandomWalkProposal(amh.Normal(0.01, 0.8))])
    elif form == "mixed_d40":
        d = 40
        target = amh.MvNormalTarget(np.linspace(0.5, 1.5, d), make_spd(d, 4, 0.2, 2.0))
        laws = _all_families(amh, d)
        props = []
        for i, law in enumerate(laws):
            if i % 3 == 0:
                props.append(amh.StaticProposal(law))
            elif i % 3 == 1:
                props.append(amh.RandomWalkProposal(amh.Normal(0.0, 0.05 + 0.01 * i)))
            else:
                props.append(amh.SymmetricRandomWalkProposal(amh.Uniform(-0.1, 0.1)))
        spl = amh.MetropolisHastings(props)
    else:
        target = amh.IIDNormalTarget(np.array([1.0]))       # README m2 shape: logpdf(Normal(x[1], x[2]), 1.0)
        spl = amh.MetropolisHastings(amh.StaticProposal([amh.Normal(0, 1), amh.InverseGamma(2, 3)]))
    rg, ro = _pair(amh, cuda, oracle, target, spl, n, _seeds(n, 60), init)
    _assert_same_state(rg, ro)
    x0 = rg.state()["x"]
    assert np.all(np.isfinite(x0))
    for k, spl_ in [(1, 1), (9, 4), (40, 0)]:
        rg.steps(k, steps_per_launch=spl_)
        ro.steps(k)
        _assert_same_state(rg, ro)
    acc = rg.state()["naccept"].sum() / (n * 50)
    assert 0.0 < acc < 1.0


def test_component_proposals_sample_and_stretch_init_bit_exact(amh, cuda, oracle):
    """`sample` over an array of proposals (schedule + save epilogue of K1C), and the ensemble's initial draw from
    `StretchProposal([InverseGamma(2,3), Normal(0,1)])` (test/emcee.jl:19)"""
    target = amh.NormalInverseGammaToy()
    spl = amh.MetropolisHastings(dict(s=amh.StaticProposal(amh.InverseGamma(2, 3)), m=amh.RandomWalkProposal(amh.Normal(0, 0.8))))
    chains = [amh.sample(np.random.default_rng(3), target, spl, amh.MCMCThreads(), 40, 300, chain_type=amh.Chains, engine=e,
                         discard_initial=7, thinning=3) for e in (cuda, oracle)]
    assert np.array_equal(chains[0].value, chains[1].value) and np.array_equal(chains[0].accepted, chains[1].accepted)
    ens = amh.Ensemble(100, amh.StretchProposal([amh.InverseGamma(2, 3), amh.Normal(0, 1)]))
    rg, ro = _pair(amh, cuda, oracle, target, ens, 300, _seeds(3, 8))
    _assert_same_state(rg, ro)
    assert np.all(rg.state()["x"][0] > 0)
    rg.steps(15); ro.steps(15)
    _assert_same_state(rg, ro)


@pytest.mark.parametrize("cfg,fwd", [("512", None), ("768", None), ("1024", None), ("512", "0"), ("512", "40")])
@pytest.mark.parametrize("levels", [None, 3])
def test_stretch_level_schedule_configs_and_overflow_bucket(amh, cuda, oracle, monkeypatch, cfg, fwd, levels):
    """K2F: every CTA size of the dataflow sweep, no / too few shared-memory forwarding slots, and the ordered overflow
    bucket (forced by capping the number of parallel levels at 3) reproduce the sequential sweep of emcee.jl:39-58
    bit for bit"""
    monkeypatch.setenv("AMH_STRETCH_BLOCK", cfg)
    if fwd is not None:
        monkeypatch.setenv("AMH_STRETCH_FWD", fwd)
    if levels is not None:
        monkeypatch.setenv("AMH_STRETCH_LEVELS", str(levels))
    target = amh.RosenbrockTarget(10)
    nw, ne = 333, 3
    spl = amh.Ensemble(nw, amh.StretchProposal(amh.MvNormal(np.zeros(10), amh.I)))
    rg, ro = _pair(amh, cuda, oracle, target, spl, nw * ne, _seeds(ne, 90))
    for k, spl_ in [(1, 1), (6, 4), (21, 0)]:
        rg.steps(k, steps_per_launch=spl_)
        ro.steps(k)
        _assert_same_state(rg, ro)
    out_g, acc_g, _ = rg.sample(5, 2, 3)
    out_o, acc_o, _ = ro.sample(5, 2, 3)
    assert np.array_equal(out_g, out_o) and np.array_equal(acc_g, acc_o)


@pytest.mark.parametrize("d,nw,ne", [(10, 333, 3), (3, 256, 2), (16, 130, 2), (12, 301, 2), (7, 129, 2), (24, 70, 2)])
@pytest.mark.parametrize("win,fwd,levels", [(None, None, None), ("7", None, None), ("100", "0", None), (None, "20", "4"),
                                              ("33", "5", "3")])
def test_stretch_resident_cluster_sweep_paths(amh, cuda, oracle, monkeypatch, d, nw, ne, win, fwd, levels):
    """K2R (amh_launch_stretch_res.cuh): the ensemble resident in the shared memory of a 2-CTA cluster, updated in place.
    Forced on for small ensembles (AMH_STRETCH_RES=1) and driven through every hand-off path: many small level-0
    windows, no / too few forwarding slots (remote version flag + distributed-shared-memory read), the ordered overflow
    bucket, odd walker counts and an odd dimension (scalar st.async) -- each reproduces the sequential sweep of
    emcee.jl:39-58 bit for bit"""
    monkeypatch.setenv("AMH_STRETCH_RES", "1")
    for k, v in (("AMH_STRETCH_WIN", win), ("AMH_STRETCH_FWD", fwd), ("AMH_STRETCH_LEVELS", levels)):
        if v is not None:
            monkeypatch.setenv(k, v)
    target = amh.RosenbrockTarget(d)
    spl = amh.Ensemble(nw, amh.StretchProposal(amh.MvNormal(np.zeros(d), amh.I)))
    rg, ro = _pair(amh, cuda, oracle, target, spl, nw * ne, _seeds(ne, 91))
    for k, spl_ in [(1, 1), (6, 4), (21, 0)]:
        rg.steps(k, steps_per_launch=spl_)
        ro.steps(k)
        _assert_same_state(rg, ro)
    out_g, acc_g, _ = rg.sample(5, 2, 3)
    out_o, acc_o, _ = ro.sample(5, 2, 3)
    assert np.array_equal(out_g, out_o) and np.array_equal(acc_g, acc_o)


def test_sample_on_four_streams_equals_one_stream_and_the_oracle(amh, cuda, oracle):
    d, n = 8, 4096 + 37
    Sigma = make_spd(d, seed=2)
    model = amh.DensityModel(amh.MvNormalTarget(None, Sigma))
    spl = amh.RWMH(amh.MvNormal(np.zeros(d), (2.38 ** 2 / d) * Sigma))
    init = np.random.default_rng(5).normal(size=(d, n))
    kw = dict(seed=11, thinning=7, initial_params=init, chain_type=amh.Chains)
    one = amh.sample(model, spl, amh.MCMCB200(device=0), 5, n, **kw)
    pout, pacc = cuda.pinned_empty((5, d + 1, n)), cuda.pinned_empty((5, n), dtype=np.uint8)
    four = amh.sample(model, spl, amh.MCMCB200(device=0, streams=4), 5, n, out=(pout, pacc), **kw)
    ref = amh.sample(model, spl, amh.MCMCB200(device=0), 5, n, engine=oracle, **kw)
    assert four.info["streams"] == 4
    assert np.array_equal(one.value, four.value) and np.array_equal(one.accepted, four.accepted)
    assert np.array_equal(four.value, ref.value)


def test_stretch_plans_made_ahead_and_mispredictions(amh, cuda, oracle):
    """K2F computes the plan of the NEXT launch on a second stream while the sweeps of the current one run; equal
    consecutive launches use it, a different length or a state reset must discard it (amh_launch_stretch.cu)."""
    d, nw, ne = 6, 1536, 3
    target = amh.RosenbrockTarget(d)
    spl = amh.Ensemble(nw, amh.StretchProposal(amh.MvNormal(np.zeros(d), amh.I)))
    rg, ro = _pair(amh, cuda, oracle, target, spl, nw * ne, _seeds(ne, 21))
    for k in (8, 8, 8, 3, 8, 8):                       # predicted, predicted, mispredicted (3), mispredicted (8), predicted
        rg.steps(k, steps_per_launch=k)
        ro.steps(k)
    _assert_same_state(rg, ro)
    st = ro.state()                                    # resume from an earlier point: the step counter jumps back
    rg.steps(5, steps_per_launch=5); ro.steps(5)
    for r in (rg, ro):
        r.set_state(st)
    rg.steps(5, steps_per_launch=5); ro.steps(5)
    rg.steps(5, steps_per_launch=5); ro.steps(5)
    _assert_same_state(rg, ro)


def test_ram_warp_redo_path_with_ieee_operators_is_bit_exact(amh, cuda, oracle, monkeypatch):
    """K4W runs its Givens sweeps speculatively with branch-free sqrt / division sequences and redoes a sweep with the
    IEEE operators from the last good factor if an operand left their exponent range.  That never happens with sane
    inputs, so the switch forces it on every step: reload, recomputed v, slow sweep -- same bits as the oracle."""
    monkeypatch.setenv("AMH_RAMW_FORCE_REDO", "1")
    for d in (20, 64):
        Sigma = make_spd(d, seed=d, lo=0.01, hi=1.0)
        target = amh.MvNormalTarget(None, Sigma)
        spl = amh.RobustAdaptiveMetropolis(eigenvalue_lower_bound=0.05, eigenvalue_upper_bound=2.0) if d == 20 else amh.RobustAdaptiveMetropolis()
        n = 300
        rg, ro = _pair(amh, cuda, oracle, target, spl, n, _seeds(n, d), np.zeros((d, n)))
        for k in (1, 12):
            rg.steps(k, warmup=True, steps_per_launch=k)
            ro.steps(k, warmup=True)
        _assert_same_state(rg, ro, S=True)


@pytest.mark.parametrize("cov", ["scalar", "diag"])
def test_symmetric_static_mh_diagonal_proposal_on_the_tensor_path(amh, cuda, oracle, cov):
    """StaticMH with issymmetric = true (Hastings term literal 0) and a diagonal proposal at a K1T16 dimension"""
    d = 16
    target = amh.MvNormalTarget(np.linspace(-0.5, 0.5, d), make_spd(d, seed=4, lo=0.5, hi=3.0))
    dist = amh.MvNormal(np.zeros(d), 1.5 * amh.I) if cov == "scalar" else [amh.Normal(0, 1.0 + 0.05 * i) for i in range(d)]
    spl = amh.MetropolisHastings(amh.SymmetricStaticProposal(dist))
    n = 777
    rg, ro = _pair(amh, cuda, oracle, target, spl, n, _seeds(n, 31))
    _assert_same_state(rg, ro)
    for k in (1, 2, 33):
        rg.steps(k, steps_per_launch=k); ro.steps(k)
        _assert_same_state(rg, ro)
