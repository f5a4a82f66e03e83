"""The multi-GPU job layer (include/amh.h amh_job_*, csrc/amh_job_impl.h): sharding, job-wide arrays, pooled summaries,
state get / set in column blocks, error propagation.  On CPU the SAME sharding code is instantiated over the oracle's
entry points ("devices" are shards), so the host logic of the N > 1 path is covered without a GPU; the `gpu` tests run
it over real devices (as many as the box has) and compare with the one-device engine and with the oracle."""
import numpy as np
import pytest

from conftest import make_spd


def _seeds(n, s=0):
    return np.random.default_rng(s).integers(0, 2 ** 64, size=n, dtype=np.uint64)


def _run_pair(eng, job, target, sampler, n, seeds, init=None):
    single = eng.run(eng.target_of(target), sampler.lower(eng, target.dim), n, seeds, init)
    multi = job.run(job.target_of(target), sampler.lower(job, target.dim), n, seeds, init)
    return single, multi


def _check_job_equals_single(eng, ngpus_list, amh):
    d = 6
    Sg = make_spd(d, 3, 0.5, 5.0)
    target = amh.MvNormalTarget(np.linspace(-1, 1, d), Sg)
    spl = amh.RWMH(amh.MvNormal(np.zeros(d), 0.4 * Sg))
    n = 37
    seeds = _seeds(n, 1)
    for k in ngpus_list:
        job = eng.job(k)
        a, b = _run_pair(eng, job, target, spl, n, seeds)
        assert sum(hi - lo for lo, hi, _ in job.shards()) == n and job.shards()[0][0] == 0
        a.steps(11); b.steps(11)
        sa, sb = a.state(), b.state()
        for key in ("x", "lp", "accepted", "naccept"):
            assert np.array_equal(sa[key], sb[key]), (k, key)
        oa, aa, ma = a.sample(9, 4, 3, 0, summary=True, chain_means=True)
        ob, ab, mb = b.sample(9, 4, 3, 0, summary=True, chain_means=True)
        assert np.array_equal(oa, ob) and np.array_equal(aa, ab)
        assert np.array_equal(ma["chain_mean"], mb["chain_mean"])
        np.testing.assert_allclose(mb["mean"], ma["mean"], rtol=1e-13, atol=1e-15)       # pooled in another order
        np.testing.assert_allclose(mb["var"], ma["var"], rtol=1e-12)
        assert mb["accept_rate"] == pytest.approx(ma["accept_rate"], rel=1e-13)
        assert (mb["n_saved"], mb["n_steps"]) == (ma["n_saved"], ma["n_steps"])
        assert b.launch_count() >= 0 and b.kernel_time_ms()[0] >= 0.0
        a.close(); b.close(); job.close()


def test_job_shards_chains_and_equals_the_single_shard_run(amh, oracle):
    _check_job_equals_single(oracle, (1, 2, 3, 5), amh)


def test_job_more_devices_than_chains_and_ensembles_are_indivisible(amh, oracle):
    target = amh.MvNormalTarget(None, np.array([[2.0, 0.3], [0.3, 1.0]]))
    job = oracle.job(4)
    a, b = _run_pair(oracle, job, target, amh.RWMH(2), 3, _seeds(3, 2))
    assert [hi - lo for lo, hi, _ in job.shards()] == [1, 1, 1, 0]                        # the fourth device idles
    a.steps(5); b.steps(5)
    assert np.array_equal(a.state()["x"], b.state()["x"])
    a.close(); b.close()
    spl = amh.Ensemble(8, amh.StretchProposal(amh.MvNormal(np.zeros(2), amh.I)))
    a, b = _run_pair(oracle, job, target, spl, 5 * 8, _seeds(5, 3))
    assert [(lo, hi) for lo, hi, _ in job.shards()] == [(0, 16), (16, 24), (24, 32), (32, 40)]   # 2 + 1 + 1 + 1 ensembles
    a.steps(7); b.steps(7)
    sa, sb = a.state(), b.state()
    for key in ("x", "lp", "accepted", "naccept"):
        assert np.array_equal(sa[key], sb[key]), key
    with pytest.raises(amh.AMHArgumentError, match="multiple of n_walkers"):
        job.run(job.target_of(target), spl.lower(job, 2), 5 * 8 + 3, _seeds(6, 3))
    job.close()


def test_job_ram_state_blocks_resume_and_failed_flags(amh, oracle):
    d, n = 5, 23
    target = amh.MvNormalTarget(None, make_spd(d, 4, 0.05, 2.0))
    spl = amh.RobustAdaptiveMetropolis(S=0.4 * np.eye(d), eigenvalue_lower_bound=0.01, eigenvalue_upper_bound=3.0)
    seeds = _seeds(n, 4)
    job = oracle.job(3)
    a, b = _run_pair(oracle, job, target, spl, n, seeds, np.zeros((d, n)))
    a.steps(15, warmup=True); b.steps(15, warmup=True)
    sa, sb = a.state(S=True), b.state(S=True)
    for key in ("x", "lp", "S", "accepted", "naccept", "logalpha", "eta", "failed"):
        assert np.array_equal(sa[key], sb[key]), key
    # resume: the job-wide state installed into a fresh job run continues bit for bit
    c = job.run(job.target_of(target), spl.lower(job, d), n, seeds, np.zeros((d, n)))
    c.set_state(sa)
    a.steps(6, warmup=True); c.steps(6, warmup=True)
    sa, sc = a.state(S=True), c.state(S=True)
    for key in ("x", "lp", "S", "accepted", "naccept", "logalpha", "eta", "step"):
        assert np.array_equal(sa[key], sc[key]), key
    # failed-downdate flags: global index of the first flagged chain
    fl = np.zeros(n, dtype=np.uint8); fl[[9, 20]] = 1
    c.set_state(dict(failed=fl))
    assert c.ram_failed()[:2] == (2, 9) and np.array_equal(c.ram_failed()[2], fl)
    a.close(); c.close(); job.close()


def test_job_errors_carry_the_device_and_the_reference_message(amh, oracle):
    job = oracle.job(2)
    A = np.array([[1.0, 0.2], [0.2, 2.0]])
    t = amh.GaussianPrecisionTarget(A)
    s2 = 0.5
    spl = amh.MALA(lambda g: amh.MvNormal((s2 / 2) * g, s2 * amh.I))
    with pytest.raises(amh.AMHStateError, match="please specify initial parameters"):      # MALA.jl:37
        job.run(job.target_of(t), spl.lower(job, 2), 4, _seeds(4))
    with pytest.raises(amh.AMHError, match="no run"):
        amh.package._capi.JobRun(job, job._target, job._sampler, 4).steps(1)
    with pytest.raises(amh.AMHArgumentError):
        oracle.job(0)
    job.close()


def test_sample_with_ngpus_equals_serial_sample_incl_resume_and_source_target(amh, oracle):
    target = amh.MvNormalTarget(None, np.array([[2.0, 0.3], [0.3, 1.0]]))
    ref = amh.sample(np.random.default_rng(3), target, amh.RWMH(2), amh.MCMCSerial(), 18, 7, chain_type=amh.Chains, engine=oracle)
    a = amh.sample(np.random.default_rng(3), target, amh.RWMH(2), amh.MCMCB200(ngpus=3), 10, 7, chain_type=amh.Chains, engine=oracle,
                   save_state=True)
    b = amh.sample(np.random.default_rng(5), target, amh.RWMH(2), amh.MCMCB200(ngpus=2), 8, 7, chain_type=amh.Chains, engine=oracle,
                   initial_state=a.info["state"])
    assert np.array_equal(np.concatenate([a.value, b.value]), ref.value)
    src = """
    AMH_TARGET double amh_user_logdensity(const double* x, int dim, const double* data, long long ndata) {
        double q = 0.0;
        for (int i = 0; i < dim; ++i) q = fma(x[i] - data[i], x[i] - data[i], q);
        return -0.5 * q;
    }"""
    st = amh.SourceTarget(3, src, data=[1.0, -2.0, 0.5])
    r1 = amh.sample(np.random.default_rng(8), st, amh.RWMH(3), amh.MCMCSerial(), 12, 5, chain_type=amh.Chains, engine=oracle)
    r2 = amh.sample(np.random.default_rng(8), st, amh.RWMH(3), amh.MCMCB200(ngpus=2), 12, 5, chain_type=amh.Chains, engine=oracle)
    assert np.array_equal(r1.value, r2.value)


# ------------------------------------------------------------------------------ on real devices
@pytest.mark.gpu
def test_job_on_the_devices_of_this_box_equals_one_device_and_the_oracle(amh, cuda, oracle):
    import torch
    ndev = torch.cuda.device_count()
    _check_job_equals_single(cuda, sorted({1, min(2, ndev), ndev}), amh)
    # C2 shape through sample(): job == oracle bit for bit, target broadcast mode reported
    d, n = 32, 4096
    Sg = make_spd(d, 32)
    target = amh.MvNormalTarget(None, Sg)
    spl = amh.RWMH(amh.MvNormal(np.zeros(d), (2.38 ** 2 / d) * Sg))
    ch = amh.sample(np.random.default_rng(1), target, spl, amh.MCMCB200(ngpus=ndev), 3, n, thinning=40, chain_type=amh.Chains)
    ro = amh.sample(np.random.default_rng(1), target, spl, amh.MCMCSerial(), 3, n, thinning=40, chain_type=amh.Chains, engine=oracle)
    assert np.array_equal(ch.value, ro.value) and np.array_equal(ch.accepted, ro.accepted)
    job = amh.sampling._job_for(cuda, ndev)
    mode, ms, init_ms = job.broadcast_info()
    assert mode in ("nccl", "peer", "h2d") and (ndev > 1 or mode == "h2d")


@pytest.mark.parametrize("spin_us", ["0", "300", None])
def test_job_worker_threads_spin_then_sleep_hand_off(amh, oracle, monkeypatch, spin_us):
    """The job's worker threads poll for the next phase for AMH_JOB_SPIN_US and then sleep (csrc/amh_job_impl.h): many short
    phases with pauses on both sides of that window must neither lose a wake-up nor change a result (0 = sleep at once,
    300 us = both regimes inside one test, default = mostly polling)."""
    import time
    if spin_us is not None:
        monkeypatch.setenv("AMH_JOB_SPIN_US", spin_us)
    target = amh.MvNormalTarget(None, np.array([[2.0, 0.3], [0.3, 1.0]]))
    spl = amh.RWMH(2)
    n = 23
    job = oracle.job(4)                      # the workers are created with the job's first fan-out and read the variable then
    a, b = _run_pair(oracle, job, target, spl, n, _seeds(n, 9))
    rng = np.random.default_rng(0)
    for i in range(120):
        k = int(rng.integers(1, 4))
        a.steps(k); b.steps(k)
        if i % 3 == 0:
            time.sleep(float(rng.choice([0.0, 0.0002, 0.002, 0.008])))       # shorter and longer than the polling window
        if i % 10 == 9:
            sa, sb = a.state(), b.state()
            assert np.array_equal(sa["x"], sb["x"]) and np.array_equal(sa["naccept"], sb["naccept"])
    a.close(); b.close(); job.close()
