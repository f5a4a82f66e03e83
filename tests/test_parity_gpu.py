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


@pytest.mark.parametrize("d,cov", [(1, "scalar"), (2, "scalar"), (2, "diag"), (3, "full"), (8, "full"), (13, "full"),
                                   (32, "full"), (32, "scalar"), (40, "full")])
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
