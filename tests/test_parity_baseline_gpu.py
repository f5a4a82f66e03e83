"""T1 parity at the BASELINE.json sizes: the CUDA engine against the CPU oracle, bit for bit,
on the shapes the bench lines are quoted on -- not only on the small cases of test_parity_gpu.py:

  C2  RWMH, MvNormal d=32, 65 536 chains x 1 000 steps            (">= 10^3 steps x 65 536 chains identical")
  C3  stretch move, Rosenbrock d=10, 64 ensembles x 4 096 walkers, 2-CTA cluster sweep, planned-ahead launches
  C4  MALA, logistic regression d=128 x 10 000 rows (1 250 ring blocks = 125 wraps of the 10-stage TMA ring)
  C5  RAM warm-up d=64 x 4 096 chains, 16 fused steps per launch, from S = I and from an adapted S0

plus the reference semantics added in round 2 (failed-downdate flag, resume of log-alpha / eta, Welford summaries,
callback states, setparams!! keeping `accepted`)."""
import numpy as np
import pytest

from conftest import make_spd

pytestmark = pytest.mark.gpu


def _pair(cuda, oracle, target, sampler, n, seeds, init=None):
    runs = []
    for eng in (cuda, oracle):
        th = eng.target(target.kind, target.dim, target.blob())
        sh = sampler.lower(eng, target.dim)
        runs.append(eng.run(th, sh, n, seeds, init))
    return runs


def _same(rg, ro, keys=("x", "lp", "accepted", "naccept"), **kw):
    sg, so = rg.state(**kw), ro.state(**kw)
    for k in keys:
        assert np.array_equal(sg[k], so[k], equal_nan=True), f"{k} differs in {np.count_nonzero(sg[k] != so[k])} entries"
    assert sg["step"] == so["step"]
    return sg


def _seeds(n, s=0):
    return np.random.default_rng(s).integers(0, 2 ** 64, size=n, dtype=np.uint64)


def test_c2_full_size_65536_chains_x_1000_steps_bit_exact(amh, cuda, oracle):
    d, n = 32, 65536
    Sigma = make_spd(d, seed=32)
    target = amh.MvNormalTarget(None, Sigma)
    spl = amh.RWMH(amh.MvNormal(np.zeros(d), (2.38 ** 2 / d) * Sigma))
    rg, ro = _pair(cuda, oracle, target, spl, n, _seeds(n, 2026))
    _same(rg, ro)                                  # first step: 65 536 draws from the proposal
    rg.steps(1000, steps_per_launch=500)           # the bench's launch shape: 500 fused steps
    ro.steps(1000)
    sg = _same(rg, ro)
    assert 0.2 < sg["naccept"].sum() / (n * 1000) < 0.3
    # and through the sampling schedule (thinning interval = one launch), with the Welford summaries
    og, ag, smg = rg.sample(3, discard_initial=100, thinning=200, chain_means=True)
    oo, ao, smo = ro.sample(3, discard_initial=100, thinning=200, chain_means=True)
    assert np.array_equal(og, oo) and np.array_equal(ag, ao)
    for k in ("mean", "var", "chain_mean"):
        assert np.array_equal(smg[k], smo[k]), k


def test_c3_full_size_64_ensembles_x_4096_walkers_cluster_sweep_bit_exact(amh, cuda, oracle):
    d, nw, ne = 10, 4096, 64
    target = amh.RosenbrockTarget(d)
    spl = amh.Ensemble(nw, amh.StretchProposal(amh.MvNormal(np.zeros(d), amh.I)))
    rg, ro = _pair(cuda, oracle, target, spl, nw * ne, _seeds(ne, 3))
    _same(rg, ro)
    rg.steps(6, steps_per_launch=3)                # two launches: the second one runs on a plan made ahead
    ro.steps(6)
    sg = _same(rg, ro)
    assert 0.2 < sg["naccept"].sum() / (nw * ne * 6) < 0.9
    og, ag, _ = rg.sample(2, discard_initial=1, thinning=2)
    oo, ao, _ = ro.sample(2, discard_initial=1, thinning=2)
    assert np.array_equal(og, oo) and np.array_equal(ag, ao)


def test_c4_full_size_10000_rows_d128_bit_exact(amh, cuda, oracle):
    d, rows, n = 128, 10000, 24                    # 24 chains = 3 warps' worth; rows are what the ring wraps over
    rng = np.random.default_rng(128)
    X = rng.normal(size=(rows, d)) / np.sqrt(d)
    beta = rng.normal(size=d)
    y = (rng.random(rows) < 1 / (1 + np.exp(-X @ beta))).astype(float)
    target = amh.LogisticRegressionTarget(X, y, tau=10.0)
    s2 = 3.3e-2                                    # tools/bench_configs.py: acceptance ~0.57 on this posterior
    spl = amh.MALA(lambda g: amh.MvNormal((s2 / 2) * g, s2 * amh.I))
    init = 0.05 * rng.normal(size=(d, n))
    rg, ro = _pair(cuda, oracle, target, spl, n, _seeds(n, 4), init)
    _same(rg, ro, keys=("x", "lp", "grad", "accepted", "naccept"), grad=True)
    rg.steps(3, steps_per_launch=2)
    ro.steps(3)
    _same(rg, ro, keys=("x", "lp", "grad", "accepted", "naccept"), grad=True)
    og, ag, _ = rg.sample(3, discard_initial=1, thinning=2)
    oo, ao, _ = ro.sample(3, discard_initial=1, thinning=2)
    assert np.array_equal(og, oo) and np.array_equal(ag, ao)
    _same(rg, ro, keys=("x", "lp", "grad", "accepted", "naccept"), grad=True)


@pytest.mark.parametrize("start", ["identity", "adapted"])
def test_c5_ram_warmup_d64_4096_chains_16_fused_steps_bit_exact(amh, cuda, oracle, start):
    d, n = 64, 4096
    Sigma = make_spd(d, 64, 1e-4, 1.0)
    target = amh.MvNormalTarget(None, Sigma)
    S0 = None if start == "identity" else (2.38 / np.sqrt(d)) * np.linalg.cholesky(Sigma)
    spl = amh.RobustAdaptiveMetropolis() if S0 is None else amh.RobustAdaptiveMetropolis(S=S0)
    rg, ro = _pair(cuda, oracle, target, spl, n, _seeds(n, 5), np.zeros((d, n)))
    keys = ("x", "lp", "S", "accepted", "naccept", "logalpha", "eta", "failed")
    _same(rg, ro, keys=keys, S=True)
    rg.steps(32, warmup=True, steps_per_launch=16)
    ro.steps(32, warmup=True)
    sg = _same(rg, ro, keys=keys, S=True)
    rg.steps(8, warmup=False, steps_per_launch=8)
    ro.steps(8, warmup=False)
    _same(rg, ro, keys=keys, S=True)
    if start == "adapted":
        assert 0.02 < sg["naccept"].sum() / (n * 32) < 0.5
    assert rg.ram_failed()[0] == ro.ram_failed()[0] == 0


# ------------------------------------------------------------------ round-2 semantics
def test_ram_failed_downdate_is_flagged_exported_and_raised_like_the_reference(amh, cuda, oracle):
    """RAM :165-171: lowrankdowndate throws PosDefException when (v_i / A_ii)^2 > 1.  With gamma = 0 (eta = 1) and
    alpha close to 1 a rejected step asks for a downdate by almost the whole factor; rounding pushes some chains over."""
    d, n = 8, 2048
    target = amh.MvNormalTarget(None, make_spd(d, 7, 1e-6, 1e-4))       # tiny target: proposals from S = I are rejected
    spl = amh.RobustAdaptiveMetropolis(alpha=1.0 - 2.0 ** -53, gamma=0.0)
    rg, ro = _pair(cuda, oracle, target, spl, n, _seeds(n, 6), np.zeros((d, n)))
    rg.steps(4, warmup=True)
    ro.steps(4, warmup=True)
    keys = ("x", "lp", "S", "accepted", "naccept", "logalpha", "eta", "failed")
    sg = _same(rg, ro, keys=keys, S=True)
    ng, fg, flg = rg.ram_failed()
    no, fo, flo = ro.ram_failed()
    assert (ng, fg) == (no, fo) and np.array_equal(flg, flo)
    if ng == 0:
        pytest.skip("no downdate failed with this seed set (the flag path is covered by the oracle comparison above)")
    assert fg == int(np.flatnonzero(flg)[0])
    with pytest.raises(amh.PosDefException) as ei:
        amh.sample(target, spl, amh.MCMCB200(), 3, n, num_warmup=4, initial_params=np.zeros((d, n)), seed=1, chain_type=amh.Chains)
    assert ei.value.count > 0
    ch = amh.sample(target, spl, amh.MCMCB200(ignore_failed_downdates=True), 3, n, num_warmup=4, initial_params=np.zeros((d, n)),
                    seed=1, chain_type=amh.Chains)
    assert ch.value.shape == (3, d + 1, n)


def test_ram_resume_carries_logalpha_eta_and_failed(amh, cuda, oracle):
    d, n = 16, 200
    target = amh.MvNormalTarget(None, make_spd(d, 8, 0.05, 2.0))
    spl = amh.RobustAdaptiveMetropolis(S=0.3 * np.eye(d))
    seeds = _seeds(n, 9)
    ra, _ = _pair(cuda, oracle, target, spl, n, seeds, np.zeros((d, n)))
    ra.steps(20, warmup=True)
    st = ra.state(S=True)
    rb, _ = _pair(cuda, oracle, target, spl, n, seeds, np.zeros((d, n)))
    rb.set_state(st)
    sb = rb.state(S=True)
    for k in ("x", "lp", "S", "logalpha", "eta", "failed", "accepted", "naccept"):
        assert np.array_equal(st[k], sb[k]), k           # a resumed run REPORTS what the uninterrupted one reports
    ra.steps(5, warmup=False); rb.steps(5, warmup=False)
    sa, sb = ra.state(S=True), rb.state(S=True)
    for k in ("x", "lp", "S", "logalpha", "eta", "accepted", "naccept"):
        assert np.array_equal(sa[k], sb[k]), k


def test_set_params_recomputes_lp_and_keeps_accepted(amh, cuda, oracle):
    """setparams!!(model, t, params) = Transition(model, params, t.accepted)  (src/AdvancedMH.jl:151-157)"""
    d, n = 3, 64
    target = amh.MvNormalTarget(None, make_spd(d, 3, 0.5, 2.0))
    spl = amh.RWMH(amh.MvNormal(np.zeros(d), 0.5 * amh.I))
    rg, ro = _pair(cuda, oracle, target, spl, n, _seeds(n, 10))
    rg.steps(5); ro.steps(5)
    before = rg.state()
    assert 0 < before["accepted"].sum() < n
    xnew = np.random.default_rng(0).normal(size=(d, n))
    rg.set_params(xnew); ro.set_params(xnew)
    sg = _same(rg, ro)
    assert np.array_equal(sg["accepted"], before["accepted"]) and np.array_equal(sg["x"], xnew)
    assert not np.array_equal(sg["lp"], before["lp"])


def test_callback_receives_the_sampler_state_of_every_saved_sample(amh, cuda, oracle):
    """test/RobustAdaptiveMetropolis.jl:11-28,46-71: StatesExtractor records `state` per saved sample and the test checks
    eigvals(S) against the bounds at every one of them"""
    d, n, lo, hi = 2, 16, 0.9, 1.1
    target = amh.MvNormalTarget(None, np.diag([10.0, 10.0]))
    spl = amh.RobustAdaptiveMetropolis(gamma=0.51, eigenvalue_lower_bound=lo, eigenvalue_upper_bound=hi)
    seen = []
    def extractor(rng, model, sampler, sample, state, iteration):
        seen.append((iteration, state.S_diag().copy(), state.logalpha.copy(), state.eta.copy(), state.iteration, sample.copy(),
                     state.isaccept.copy()))
    N = 60
    ch = amh.sample(target, spl, amh.MCMCB200(), N, n, num_warmup=N, discard_initial=0, callback=extractor, seed=3,
                    initial_params=np.zeros((d, n)), chain_type=amh.Chains)
    assert [s[0] for s in seen] == list(range(1, N + 1))
    assert [s[4] for s in seen] == list(range(1, N + 1))            # state.iteration: 1 after the first step (RAM :211)
    diag = np.stack([s[1] for s in seen])
    assert (diag >= lo).all() and (diag <= hi).all()                 # every eigenvalue of every state within the bounds
    assert diag.max() == pytest.approx(hi, abs=0.05)                 # sigma^2 = 10: the upper bound is reached
    assert all((s[2] <= 0).all() for s in seen)
    np.testing.assert_allclose(seen[-1][3], float(N - 1) ** -0.51)   # eta of the last adaptation: iteration^-gamma
    assert np.array_equal(np.stack([s[5] for s in seen]), ch.value)
    assert np.array_equal(np.stack([s[6] for s in seen]), ch.accepted.astype(bool))
    # and the stored samples are what the plain (no-callback) schedule returns
    ch2 = amh.sample(target, spl, amh.MCMCB200(), N, n, num_warmup=N, discard_initial=0, seed=3,
                     initial_params=np.zeros((d, n)), chain_type=amh.Chains)
    assert np.array_equal(ch.value, ch2.value) and np.array_equal(ch.accepted, ch2.accepted)


# ------------------------------------------------------------------ contract versions
@pytest.mark.parametrize("cv", [1, 2])
def test_both_contract_versions_on_every_tuned_kernel_bit_exact(amh, cuda, oracle, cv):
    """the default contract is v2 (Philox4x32-7, four normals per block); v1 stays selectable per run and every kernel
    implements both: K1T16 (d = 8, 16, 24, 32; full and diagonal proposals), K1 fixed / generic dimensions, K3L, K4W, K4"""
    with amh.contract(cv):
        for d, cov in ((8, "full"), (16, "full"), (24, "full"), (32, "full"), (32, "diag"), (10, "full"), (13, "full"), (40, "full")):
            Sg = make_spd(d, seed=d)
            target = amh.MvNormalTarget(np.linspace(-1, 1, d), Sg)
            prop = amh.MvNormal(np.zeros(d), (2.38 ** 2 / d) * Sg) if cov == "full" else [amh.Normal(0, 0.5 + 0.1 * i) for i in range(d)]
            n = 777
            rg, ro = _pair(cuda, oracle, target, amh.RWMH(prop), n, _seeds(n, d))
            assert rg.contract() == ro.contract() == cv
            _same(rg, ro)
            for k, spl in ((1, 1), (9, 4), (33, 0)):
                rg.steps(k, steps_per_launch=spl); ro.steps(k)
                _same(rg, ro)
        # static MH on the tensor path (odd / even launch steps share the exponential draw between the two lanes of a chain)
        d = 16
        Sg = make_spd(d, 5, 0.5, 3.0)
        rg, ro = _pair(cuda, oracle, amh.MvNormalTarget(None, Sg), amh.MetropolisHastings(amh.StaticProposal(amh.MvNormal(np.zeros(d), 1.3 * Sg))),
                       300, _seeds(300, 6))
        rg.steps(7, steps_per_launch=3); ro.steps(7)
        _same(rg, ro)
        # K3L d = 32 / 64 / 128
        for d, rows, n in ((32, 100, 50), (64, 72, 24), (128, 203, 70)):
            rng = np.random.default_rng(d)
            X = rng.normal(size=(rows, d)) / np.sqrt(d)
            y = (rng.random(rows) < 0.5).astype(float)
            s2 = 0.02
            rg, ro = _pair(cuda, oracle, amh.LogisticRegressionTarget(X, y, tau=5.0), amh.MALA(lambda g: amh.MvNormal((s2 / 2) * g, s2 * amh.I)),
                           n, _seeds(n, 300 + d), 0.1 * rng.normal(size=(d, n)))
            rg.steps(6, steps_per_launch=2); ro.steps(6)
            _same(rg, ro, keys=("x", "lp", "grad", "accepted", "naccept"), grad=True)
        # MALA on the per-thread kernel (fixed and generic dimension), RAM on K4W and on K4
        for d in (5, 16):
            Sg = make_spd(d, 9, 0.5, 3.0)
            s2 = 0.1
            rg, ro = _pair(cuda, oracle, amh.MvNormalTarget(None, Sg), amh.MALA(lambda g: amh.MvNormal((s2 / 2) * g, s2 * amh.I)), 200, _seeds(200, 7),
                           np.zeros((d, 200)))
            rg.steps(9); ro.steps(9)
            _same(rg, ro, keys=("x", "lp", "grad", "accepted", "naccept"), grad=True)
        for d in (2, 20, 33, 64):
            Sg = make_spd(d, 11, 0.05, 2.0)
            rg, ro = _pair(cuda, oracle, amh.MvNormalTarget(None, Sg), amh.RobustAdaptiveMetropolis(S=0.3 * np.eye(d)), 150, _seeds(150, 8))
            rg.steps(12, warmup=True, steps_per_launch=5); ro.steps(12, warmup=True)
            _same(rg, ro, keys=("x", "lp", "S", "accepted", "naccept", "logalpha", "eta"), S=True)
        # arrays of univariate laws: Normal components read the step's Box-Muller slots
        comp = amh.MetropolisHastings(amh.StaticProposal([amh.Normal(0, 1), amh.InverseGamma(2, 3), amh.Normal(1, 2)]))
        rg, ro = _pair(cuda, oracle, amh.MvNormalTarget(np.array([0.0, 1.5, 1.0]), np.eye(3)), comp, 100, _seeds(100, 9))
        rg.steps(15); ro.steps(15)
        _same(rg, ro)


# ------------------------------------------------------------------ K3T: the opt-in split-bf16 tcgen05 path (tolerance, not bit-exact)
def _logistic(rows, d, seed):
    rng = np.random.default_rng(seed)
    X = rng.normal(size=(rows, d)) / np.sqrt(d)
    y = (rng.random(rows) < 1 / (1 + np.exp(-X @ rng.normal(size=d)))).astype(float)
    return X, y


@pytest.mark.parametrize("rows,n", [(1000, 200), (64, 128), (10000, 300)])
def test_mala_tensor_path_one_step_within_stated_tolerance(amh, cuda, oracle, rows, n):
    """precision = bf16x2 (csrc/amh_launch_mala_tensor.cu): eta = X c and X'r as split-bf16 tcgen05 GEMMs with fp32
    accumulation, everything else fp64.  STATED TOLERANCE against the fp64 oracle, one step from the same state:
      * the candidate is fp64 and identical, so accepted chains hold bit-identical x;
      * log-density of the new state: |lp - lp_oracle| <= 2e-3 + 2e-6 |lp|   (observed ~1e-4);
      * gradient: max_j |g_j - g_j_oracle| <= 2e-3 * max(1, max_j |g_j|);
      * accept / reject decisions agree except when |log-alpha + e| is inside that tolerance (<= 1 % of the chains here)."""
    d = 128
    X, y = _logistic(rows, d, rows)
    target = amh.LogisticRegressionTarget(X, y, tau=10.0)
    s2 = 3.3e-2 if rows >= 5000 else 0.2
    spl = amh.MALA(lambda g: amh.MvNormal((s2 / 2) * g, s2 * amh.I))
    init = 0.05 * np.random.default_rng(1).normal(size=(d, n))
    seeds = _seeds(n, 11)
    with amh.precision("bf16x2"):
        rg = cuda.run(cuda.target_of(target), spl.lower(cuda, d), n, seeds, init)
    ro = oracle.run(oracle.target_of(target), spl.lower(oracle, d), n, seeds, init)
    s0g, s0o = rg.state(grad=True), ro.state(grad=True)
    for k in ("x", "lp", "grad"):
        assert np.array_equal(s0g[k], s0o[k]), k              # the first step (initial lp / gradient) is the fp64 init kernel
    rg.steps(1); ro.steps(1)
    sg, so = rg.state(grad=True), ro.state(grad=True)
    agree = sg["accepted"] == so["accepted"]
    assert agree.mean() >= 0.99, f"accept decisions differ for {(~agree).sum()} of {n} chains"
    both = agree & (sg["accepted"] == 1)
    assert both.sum() > 0.3 * n
    assert np.array_equal(sg["x"][:, both], so["x"][:, both])                     # fp64 candidates, identical
    dlp = np.abs(sg["lp"][agree] - so["lp"][agree])
    assert dlp.max() <= 2e-3 + 2e-6 * np.abs(so["lp"]).max(), dlp.max()
    gmax = max(1.0, np.abs(so["grad"]).max())
    assert np.abs(sg["grad"][:, agree] - so["grad"][:, agree]).max() <= 2e-3 * gmax
    rg.close(); ro.close()


def test_mala_tensor_path_long_run_statistics_and_api(amh, cuda, oracle):
    """many steps: the chains decouple from the fp64 ones at the first knife-edge decision, so the comparison is
    statistical -- acceptance rate and posterior means of the fp64 engine (K3L) and of the tensor path agree"""
    d, rows, n = 128, 2000, 1024
    X, y = _logistic(rows, d, 5)
    target = amh.LogisticRegressionTarget(X, y, tau=10.0)
    s2 = 0.12
    spl = amh.MALA(lambda g: amh.MvNormal((s2 / 2) * g, s2 * amh.I))
    init = np.zeros((d, n))
    kw = dict(initial_params=init, discard_initial=60, thinning=5, seed=3, chain_type=amh.Chains)
    a = amh.sample(target, spl, amh.MCMCB200(), 8, n, **kw)
    b = amh.sample(target, spl, amh.MCMCB200(dtype="bf16x2"), 8, n, **kw)
    assert abs(a.accepted[1:].mean() - b.accepted[1:].mean()) < 0.02
    ma, mb = a.value[1:, :d, :].mean(axis=(0, 2)), b.value[1:, :d, :].mean(axis=(0, 2))
    sd = a.value[1:, :d, :].std(axis=(0, 2))
    assert np.abs(ma - mb).max() < 0.1 * sd.max()                                  # 7 x 1024 draws: standard error ~0.012 sd
    assert np.abs(a.value[1:, d, :].mean() - b.value[1:, d, :].mean()) < 0.5      # mean log-density (std of lp ~ 8)
    # only MALA x logistic x dim 128 has this path; everything else says so instead of silently running fp64
    with pytest.raises(amh.AMHArgumentError, match="bf16x2"):
        amh.sample(amh.MvNormalTarget(None, np.eye(4)), amh.RWMH(4), amh.MCMCB200(dtype="bf16x2"), 3, 8)
