"""Host-side mirror of the reference interface (constructors, `sample` keywords, chain types, errors),
exercised on the CPU oracle engine so that it runs without a GPU.  Mirrors test/runtests.jl:37-54,
112-213, 288-303 of the reference."""
import math

import numpy as np
import pytest

from conftest import make_spd


@pytest.fixture(scope="module")
def model(amh):
    data = np.random.default_rng(1234).normal(0, 1, 30)
    return amh.DensityModel(amh.IIDNormalTarget(data))


def test_constructor_lowering(amh, oracle):
    for spl, cov in [(amh.RWMH(3), "scalar"), (amh.StaticMH([amh.Normal(0, 1)] * 3), "diag"),
                     (amh.RWMH(amh.MvNormal(np.zeros(3), make_spd(3, 1))), "full")]:
        assert spl.dim == 3
        spl.lower(oracle, 3).close()
    assert amh.RWMH(2).proposal.issymmetric is False                      # proposal.jl:18 default
    assert amh.SymmetricRandomWalkProposal(amh.Normal()).issymmetric is True
    with pytest.raises(ValueError, match="dimension"):
        amh.RWMH(3).lower(oracle, 2)
    with pytest.raises(ValueError, match="function-valued"):
        amh.MetropolisHastings(amh.RandomWalkProposal(lambda x: amh.Normal(x, 1)))
    with pytest.raises(ValueError):
        amh.MetropolisHastings([amh.Normal(0, 1)])                        # bare containers are host-only
    with pytest.raises(ValueError, match="catalogue"):
        amh.DensityModel(lambda x: -0.5 * x @ x)                         # closures cannot run on the device


def test_mvnormal_forms(amh):
    I = amh.I
    a = amh.MvNormal(np.zeros(4), I)
    assert a.kind == "scalar" and a.scale[0] == 1.0 and a.zero_mean
    b = amh.MvNormal(np.ones(2), 0.25 * I)
    assert b.scale[0] == 0.5 and not b.zero_mean
    c = amh.MvNormal(np.zeros(2), np.array([4.0, 9.0]))
    assert c.kind == "diag" and np.array_equal(c.scale, [2.0, 3.0])
    S = make_spd(3, 2)
    d = amh.MvNormal(np.zeros(3), S)
    L = np.zeros((3, 3)); L[np.tril_indices(3)] = d.scale
    assert np.allclose(L @ L.T, S)
    with pytest.raises(ValueError):
        amh.MvNormal(np.zeros(2), -1.0 * I)


def test_mala_closure_probing(amh, oracle):
    s2 = 0.3
    spl = amh.MALA(lambda g: amh.MvNormal((s2 / 2) * g, s2 * amh.I))
    assert spl.probe(3) == (pytest.approx(s2), pytest.approx(s2 / 2))
    # MALA wraps a RandomWalkProposal too (MALA.jl:8-11)
    spl2 = amh.MALA(amh.RandomWalkProposal(lambda g: amh.MvNormal(0.1 * g, s2 * amh.I)))
    assert spl2.probe(2) == (pytest.approx(s2), pytest.approx(0.1))
    with pytest.raises(ValueError):
        amh.MALA(lambda g: amh.MvNormal(g * g, s2 * amh.I)).probe(2)        # not linear in g
    with pytest.raises(ValueError):
        amh.MALA(lambda g: amh.MvNormal(0.1 * g, np.array([1.0, 2.0]))).probe(2)   # not sigma2 * I
    with pytest.raises(ValueError):
        amh.MALA(amh.Normal())


def test_sample_single_chain_and_chain_types(amh, oracle, model):
    ch = amh.sample(model, amh.RWMH(2), 50, engine=oracle, seed=1)
    assert len(ch) == 50 and ch[3].params.shape == (2,) and isinstance(ch[3].lp, float)
    c2 = amh.sample(model, amh.RWMH(2), 50, engine=oracle, seed=1, chain_type=amh.Chains)
    assert c2.value.shape == (50, 3, 1) and c2.names == ["param_1", "param_2", "lp"]
    assert np.array_equal(c2.value[3, :2, 0], ch[3].params)             # same seed, same chain
    sa = amh.sample(model, amh.RWMH(2), 50, engine=oracle, seed=1, chain_type=amh.StructArray, param_names=["a", "b"])
    assert set(sa.keys()) == {"a", "b", "lp"} and np.array_equal(sa.a, c2["param_1"][:, 0])
    nt = amh.sample(model, amh.RWMH(2), 5, engine=oracle, seed=1, chain_type="namedtuples")
    assert list(nt[0].keys()) == ["param_1", "param_2", "lp"]
    with pytest.raises(ValueError, match="param_names"):
        amh.sample(model, amh.RWMH(2), 5, engine=oracle, param_names=["a"], chain_type=amh.Chains)


def test_sample_multi_chain_matches_single_chains(amh, oracle, model):
    """chains are independent streams: chain c of a 4-chain call equals a 1-chain run seeded with seeds[c]"""
    rng = np.random.default_rng(5)
    ch = amh.sample(rng, model, amh.RWMH(2), amh.MCMCThreads(), 40, 4, chain_type=amh.Chains, engine=oracle,
                    initial_params=[[0.0, 1.0]] * 4)
    assert ch.value.shape == (40, 3, 4)
    assert np.array_equal(ch.value[0, :2, :], np.tile([[0.0], [1.0]], (1, 4)))
    seeds = np.random.default_rng(5).integers(0, 2 ** 64, size=4, dtype=np.uint64)
    t = model.logdensity
    for c in range(4):
        run = oracle.run(oracle.target(t.kind, 2, t.blob()), amh.RWMH(2).lower(oracle, 2), 1, seeds[c:c + 1],
                         np.array([[0.0], [1.0]]))
        out, _, _ = run.sample(40)
        assert np.array_equal(out[:, :, 0], ch.value[:, :, c])
    with pytest.raises(ValueError, match="one entry per chain"):
        amh.sample(model, amh.RWMH(2), amh.MCMCSerial(), 10, 4, engine=oracle, initial_params=[[0.0, 1.0]] * 3)
    assert ch.array().shape == (160, 2)


def test_discard_thinning_warmup_defaults(amh, oracle):
    target = amh.MvNormalTarget(None, np.eye(2))
    ch = amh.sample(target, amh.RobustAdaptiveMetropolis(), 20, num_warmup=30, chain_type=amh.Chains, engine=oracle,
                    summary=True)
    assert ch.start == 31 and ch.info["summary"]["n_steps"] == 30 + 19       # discard_initial defaults to num_warmup
    ch = amh.sample(target, amh.RWMH(2), 7, discard_initial=3, thinning=5, chain_type=amh.Chains, engine=oracle,
                    summary=True)
    assert list(ch.range()) == [4 + 5 * i for i in range(7)] and ch.info["summary"]["n_steps"] == 3 + 30


def test_mala_errors_match_reference(amh, oracle):
    target = amh.GaussianPrecisionTarget(np.eye(2))
    spl = amh.MALA(lambda g: amh.MvNormal(0.25 * g, 0.5 * amh.I))
    with pytest.raises(amh.AMHStateError, match="please specify initial parameters"):      # MALA.jl:37
        amh.sample(target, spl, 10, engine=oracle)
    with pytest.raises(amh.AMHArgumentError, match="gradient"):                            # MALA.jl:42-52
        amh.sample(amh.NormalInverseGammaToy(), spl, 10, engine=oracle, initial_params=np.ones(2))
    st = amh.sample(target, spl, 10, engine=oracle, initial_params=np.ones(2))
    assert np.array_equal(st[0].params, np.ones(2))


def test_getparams_setparams_roundtrip(amh, oracle):
    """test/runtests.jl:37-54: setparams!! recomputes lp (and the gradient for MALA); RAM keeps logprob"""
    target = amh.GaussianPrecisionTarget(np.array([[2.0, 0.3], [0.3, 1.0]]))
    th = oracle.target(target.kind, 2, target.blob())
    x0 = np.ones((2, 3)); x1 = np.array([[0.5, -1.0, 2.0], [0.1, 0.2, 0.3]])
    run = oracle.run(th, amh.MALA(lambda g: amh.MvNormal(0.05 * g, 0.1 * amh.I)).lower(oracle, 2), 3, np.arange(3, dtype=np.uint64), x0)
    run.set_params(x1)
    st = run.state(grad=True)
    A = target.A
    assert np.array_equal(st["x"], x1)
    assert np.allclose(st["lp"], [-0.5 * x1[:, c] @ A @ x1[:, c] for c in range(3)])
    assert np.allclose(st["grad"], -A @ x1)
    run = oracle.run(th, amh.RobustAdaptiveMetropolis().lower(oracle, 2), 3, np.arange(3, dtype=np.uint64), x0)
    lp0 = run.state()["lp"].copy()
    run.set_params(x1)
    assert np.array_equal(run.state()["lp"], lp0)                          # RAM :117-121


def test_ensemble_layout_and_errors(amh, oracle):
    target = amh.RosenbrockTarget(4)
    spl = amh.Ensemble(16, amh.StretchProposal(amh.MvNormal(np.zeros(4), amh.I)))
    ch = amh.sample(target, spl, 12, chain_type=amh.Chains, engine=oracle)
    assert ch.value.shape == (12, 5, 16)        # nsamples x (d+1) x n_walkers  (ext/AdvancedMHMCMCChainsExt.jl:93-106)
    tv = amh.sample(target, spl, 12, engine=oracle)
    assert len(tv[0]) == 16 and tv[0][0].params.shape == (4,)
    with pytest.raises(ValueError):
        amh.Ensemble(16, amh.MvNormal(np.zeros(4), amh.I))
    with pytest.raises(amh.AMHArgumentError, match="n_walkers"):
        amh.sample(target, amh.Ensemble(1, amh.StretchProposal(amh.MvNormal(np.zeros(4), amh.I))), 3, engine=oracle)
    ch2 = amh.sample(target, spl, amh.MCMCSerial(), 5, 3, chain_type=amh.Chains, engine=oracle)
    assert ch2.value.shape == (5, 5, 48)


def test_callback_and_summary_only(amh, oracle, model):
    seen = []
    amh.sample(model, amh.RWMH(2), 6, engine=oracle, callback=lambda rng, m, s, sample, state, i: seen.append(i))
    assert seen == [1, 2, 3, 4, 5, 6]
    info = amh.sample(model, amh.RWMH(2), amh.MCMCSerial(), 100, 8, engine=oracle, store=False, summary=True)
    assert info["summary"]["n_saved"] == 100 and info["summary"]["mean"].shape == (2,)
    assert 0 < info["summary"]["accept_rate"] < 1


def test_shard_bounds_partition(amh):
    for n, w in [(10, 3), (8, 8), (5, 8), (262144, 8), (64, 8)]:
        parts = [amh.shard_bounds(n, r, w) for r in range(w)]
        assert parts[0][0] == 0 and parts[-1][1] == n
        assert all(parts[i][1] == parts[i + 1][0] for i in range(w - 1))
        sizes = [b - a for a, b in parts]
        assert max(sizes) - min(sizes) <= 1


def test_caller_buffers_pinned_alloc_and_strided_init(amh, oracle):
    """`out=` caller buffers, Engine.pinned_empty (amh_host_alloc; malloc-backed in the oracle) and a column block of
    a larger initial_params matrix (init_ld) -- the pieces bench.py's e2e path relies on"""
    target = amh.MvNormalTarget(None, np.array([[2.0, 0.3], [0.3, 1.0]]))
    spl = amh.RWMH(2)
    N, n = 6, 5
    pout = oracle.pinned_empty((N, 3, n)); pacc = oracle.pinned_empty((N, n), dtype=np.uint8)
    big = np.random.default_rng(0).normal(size=(2, 11))
    ch = amh.sample(np.random.default_rng(1), target, spl, amh.MCMCSerial(), N, n, chain_type=amh.Chains, engine=oracle,
                    initial_params=np.ascontiguousarray(big[:, 3:8]), out=(pout, pacc))
    assert ch.value is pout and ch.accepted is pacc
    ref = amh.sample(np.random.default_rng(1), target, spl, amh.MCMCSerial(), N, n, chain_type=amh.Chains, engine=oracle,
                     initial_params=[big[:, 3 + c] for c in range(n)])
    assert np.array_equal(ch.value, ref.value) and np.array_equal(ch.accepted, ref.accepted)
    # strided view passed straight to the engine
    th = oracle.target(target.kind, 2, target.blob()); sh = spl.lower(oracle, 2)
    seeds = np.arange(n, dtype=np.uint64)
    r1 = oracle.run(th, sh, n, seeds, big[:, 3:8])
    r2 = oracle.run(th, sh, n, seeds, np.ascontiguousarray(big[:, 3:8]))
    assert np.array_equal(r1.state()["x"], r2.state()["x"]) and np.array_equal(r1.state()["lp"], r2.state()["lp"])
    with pytest.raises(amh.AMHArgumentError):
        r1.sample(3, out=np.empty((3, 3, n + 1)))


@pytest.mark.parametrize("kind", ["rwmh", "static", "mala", "ram", "stretch"])
def test_initial_state_resume_continues_bit_for_bit(amh, oracle, kind):
    """`initial_state=` / `save_state=`: two calls == one long call, for every sampler's state layout (host logic + oracle)"""
    S = np.array([[2.0, 0.3, 0.0], [0.3, 1.0, 0.2], [0.0, 0.2, 0.7]])
    target = amh.MvNormalTarget(None, S)
    kw, nch, wu = {}, 6, 0
    if kind == "rwmh":
        spl = amh.RWMH(amh.MvNormal(np.zeros(3), 0.5 * S))
    elif kind == "static":
        spl = amh.MetropolisHastings(amh.StaticProposal(amh.MvNormal(np.full(3, 0.1), 1.5 * S)))
    elif kind == "mala":
        target = amh.GaussianPrecisionTarget(np.linalg.inv(S))
        spl = amh.MALA(lambda g: amh.MvNormal(0.1 * g, 0.2 * amh.I))
        kw = dict(initial_params=[np.ones(3)] * nch)
    elif kind == "ram":
        spl, wu = amh.RobustAdaptiveMetropolis(), 14
    else:
        spl, nch = amh.Ensemble(9, amh.StretchProposal(amh.MvNormal(np.zeros(3), amh.I))), 2
    common = dict(chain_type=amh.Chains, engine=oracle, num_warmup=wu, discard_initial=0)
    full = amh.sample(np.random.default_rng(9), target, spl, amh.MCMCThreads(), 30, nch, **common, **kw)
    a = amh.sample(np.random.default_rng(9), target, spl, amh.MCMCThreads(), 10, nch, save_state=True, **common, **kw)
    st = a.info["state"]
    assert st["step"] == 9 and st["x"].shape[0] == 3
    b = amh.sample(np.random.default_rng(77), target, spl, amh.MCMCThreads(), 20, nch, initial_state=st, **common)
    assert np.array_equal(np.concatenate([a.value, b.value]), full.value)
    assert np.array_equal(np.concatenate([a.accepted, b.accepted]), full.accepted)
    with pytest.raises(ValueError):
        amh.sample(target, spl, amh.MCMCThreads(), 5, nch + 1, initial_state=st, **common)


def test_component_proposal_lowering_and_errors(amh, oracle):
    """arrays of univariate laws / arrays of proposals (README.md:104-133): lowering, dimension and parameter checks"""
    K = amh.package._capi
    s = amh.MetropolisHastings(amh.StaticProposal([amh.Normal(0, 1), amh.InverseGamma(2, 3)]))
    assert s.dim == 2 and s.kind == K.SAMPLER_STATIC and [c[0] for c in s.components] == [1, 2]
    # an all-Normal array keeps the diagonal-Gaussian path
    assert amh.MetropolisHastings(amh.StaticProposal([amh.Normal(0, 1), amh.Normal(0, 2)])).components is None
    m = amh.MetropolisHastings([amh.StaticProposal(amh.Normal(0, 1)), amh.SymmetricRandomWalkProposal(amh.Uniform(-1, 1))])
    assert m.kind == K.SAMPLER_MIXED and m.components[1][4:] == (True, True) and m.components[0][4:] == (False, False)
    nt = amh.MetropolisHastings(dict(a=amh.StaticProposal(amh.Normal(0, 1)), b=amh.StaticProposal(amh.InverseGamma(2, 3))))
    assert nt.names == ["a", "b"]
    h = m.lower(oracle, 2); h.close()
    with pytest.raises(ValueError):
        m.lower(oracle, 3)
    with pytest.raises(ValueError):
        amh.MetropolisHastings([amh.StaticProposal(amh.MvNormal(np.zeros(2), amh.I))])       # not univariate
    with pytest.raises(ValueError):
        amh.InverseGamma(-1, 3)
    with pytest.raises(ValueError):
        oracle.sampler(kind=K.SAMPLER_STATIC, dim=1, cov_kind=K.COV_COMPONENTS, components=[(2, -1.0, 3.0, 0.0)])
    with pytest.raises(ValueError):
        oracle.sampler(kind=K.SAMPLER_STATIC, dim=1, cov_kind=K.COV_COMPONENTS, components=[(99, 1.0, 3.0, 0.0)])
    with pytest.raises(ValueError):
        oracle.sampler(kind=K.SAMPLER_MIXED, dim=2, components=[(1, 0.0, 1.0, 0.0)])
    # README c2: sample(m2, MetropolisHastings(p2), 100; chain_type=Vector{NamedTuple})
    c2 = amh.sample(amh.DensityModel(amh.IIDNormalTarget(np.array([1.0]))), s, 100, chain_type="namedtuples", engine=oracle,
                    param_names=["mu", "sigma"])
    assert len(c2) == 100 and set(c2[0]) == {"mu", "sigma", "lp"} and all(r["sigma"] > 0 for r in c2)


def test_stream_shards_reproduce_the_single_stream_run(amh, oracle):
    """MCMCB200(streams=k): contiguous shards of a rank's chains, one context + host thread each, filling column blocks
    of one output array (amh_run_sample_ld) -- identical to the unsharded run (global chain identity)."""
    d = 3
    model = amh.DensityModel(amh.MvNormalTarget(None, np.eye(d) * 0.5 + 0.5))
    spl = amh.RWMH(amh.MvNormal(np.zeros(d), 0.3 * amh.I))
    init = np.random.default_rng(0).normal(size=(d, 53))
    kw = dict(seed=3, thinning=3, discard_initial=5, initial_params=init, chain_type=amh.Chains, engine=oracle)
    a = amh.sample(model, spl, amh.MCMCB200(streams=1), 12, 53, **kw)
    b = amh.sample(model, spl, amh.MCMCB200(streams=4), 12, 53, **kw)
    assert b.info["streams"] == 4
    assert np.array_equal(a.value, b.value) and np.array_equal(a.accepted, b.accepted)
    ens = amh.Ensemble(8, amh.StretchProposal(amh.MvNormal(np.zeros(d), amh.I)))      # an ensemble is indivisible
    a = amh.sample(model, ens, amh.MCMCB200(streams=1), 6, 5, seed=1, chain_type=amh.Chains, engine=oracle)
    b = amh.sample(model, ens, amh.MCMCB200(streams=2), 6, 5, seed=1, chain_type=amh.Chains, engine=oracle)
    assert np.array_equal(a.value, b.value)


def test_run_sample_ld_rejects_a_leading_dimension_smaller_than_the_shard(amh, oracle):
    d, n = 2, 9
    t = amh.MvNormalTarget(None, np.eye(d))
    run = oracle.run(oracle.target(t.kind, d, t.blob()), amh.RWMH(d).lower(oracle, d), n, np.arange(n, dtype=np.uint64))
    big = np.zeros((3, d + 1, 2 * n))
    out, acc, _ = run.sample(3, out=big[:, :, n:], acc=None, summary=False)          # a block of chains of a larger array
    assert np.all(big[:, :, :n] == 0) and np.any(big[:, :, n:] != 0)
    with pytest.raises(amh.AMHArgumentError):
        run.sample(3, out=np.zeros((3, d + 1, n))[:, :, ::2], summary=False)        # not unit stride along chains


def test_python_host_points_the_job_layer_at_the_nccl_next_to_pytorch(monkeypatch):
    """_capi._point_at_bundled_nccl: without AMH_NCCL_LIB the library would dlopen the system libnccl.so.2 under its soname, and a
    later `import torch` would be handed that (older) copy and fail; the host names the wheel's copy instead, never overriding
    a caller's choice"""
    import importlib.util
    import os
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import amh_b200  # noqa: F401
    import advancedmh_jl_b200._capi as K
    monkeypatch.delenv("AMH_NCCL_LIB", raising=False)
    K._point_at_bundled_nccl()
    spec = importlib.util.find_spec("nvidia.nccl")
    if spec is not None and any(os.path.exists(os.path.join(p, "lib", "libnccl.so.2")) for p in spec.submodule_search_locations):
        assert os.environ["AMH_NCCL_LIB"].endswith(os.path.join("nvidia", "nccl", "lib", "libnccl.so.2"))
    else:
        assert "AMH_NCCL_LIB" not in os.environ
    monkeypatch.setenv("AMH_NCCL_LIB", "/somewhere/else/libnccl.so.2")
    K._point_at_bundled_nccl()
    assert os.environ["AMH_NCCL_LIB"] == "/somewhere/else/libnccl.so.2"
