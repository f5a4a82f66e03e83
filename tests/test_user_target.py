"""User-supplied targets: the log-density arrives as source text (amh_target_create_source) -- the
stand-in for DensityModel(f) with an arbitrary closure (src/AdvancedMH.jl:52-54) and for LogDensityProblems objects
(src/AdvancedMH.jl:76, MALA.jl:100-105).

CPU part: the oracle compiles the same text with g++ (contract flags) and must reproduce the catalogue targets bit for
bit when the text restates them; the NVRTC translation unit of the product compiles for sm_100a (no GPU needed).
GPU part (-m gpu): every sampler on NVRTC-compiled targets against the oracle on the same seeds, bit-exact."""
import ctypes
import os
import sys

import numpy as np
import pytest

from conftest import ROOT, make_spd

ROSENBROCK_SRC = r"""
/* -sum_{i<d-1} [b (x_{i+1} - x_i^2)^2 + (a - x_i)^2] / s  with data = [a, b, s] */
AMH_TARGET double amh_user_logdensity(const double* x, int dim, const double* data, long long ndata) {
    double acc = 0.0;
    for (int i = 0; i + 1 < dim; ++i) {
        const double t1 = fma(-x[i], x[i], x[i + 1]);
        const double t2 = data[0] - x[i];
        acc = acc + fma(data[1] * t1, t1, t2 * t2);
    }
    return -(acc / data[2]);
}
AMH_TARGET void amh_user_logdensity_and_gradient(const double* x, int dim, const double* data, long long ndata,
                                                 double* lp, double* grad) {
    *lp = amh_user_logdensity(x, dim, data, ndata);
    for (int i = 0; i < dim; ++i) grad[i] = 0.0;
    for (int i = 0; i + 1 < dim; ++i) {
        const double t1 = fma(-x[i], x[i], x[i + 1]);
        const double t2 = data[0] - x[i];
        grad[i] = grad[i] + (-4.0 * data[1] * t1 * x[i] - 2.0 * t2);
        grad[i + 1] = grad[i + 1] + 2.0 * data[1] * t1;
    }
    for (int i = 0; i < dim; ++i) grad[i] = -(grad[i] / data[2]);
}
"""

# Neal's funnel with a data-dependent likelihood: exercises the contract's exp/log and the data pointer
FUNNEL_SRC = r"""
AMH_TARGET double amh_user_logdensity(const double* x, int dim, const double* data, long long ndata) {
    const double v = x[0];                       /* log-variance ~ N(0, 3^2) */
    if (!(fabs(v) < 30.0)) return -INFINITY;
    double lp = -0.5 * (v * v / 9.0);
    const double iv = amh::exp_(-v);
    for (int i = 1; i < dim; ++i) lp = lp + (-0.5 * (x[i] * x[i] * iv) - 0.5 * v);
    for (long long k = 0; k < ndata; ++k) {       /* y_k ~ Bernoulli(sigmoid(x[1] + ... )) style term */
        const double eta = x[1] * data[k];
        lp = lp - amh::log1pexp(-eta);
    }
    return lp;
}
"""

IID_NORMAL_SRC = r"""
/* README.md:26-31: sum(logpdf.(Normal(mu, sigma), data)) on sigma >= 0, theta = (mu, sigma) */
AMH_TARGET double amh_user_logdensity(const double* x, int dim, const double* data, long long ndata) {
    const double mu = x[0], sigma = x[1];
    if (!(sigma >= 0.0)) return -INFINITY;
    const double ls = amh::log_(sigma);
    double acc = 0.0;
    for (long long i = 0; i < ndata; ++i) {
        const double z = (data[i] - mu) / sigma;
        const double t = z * z + AMH_LOG_2PI;
        acc = acc + (-0.5 * t - ls);
    }
    return acc;
}
"""

BROKEN_SRC = "AMH_TARGET double amh_user_logdensity(const double* x, int dim, const double* data, long long ndata) { return undefined_symbol(x); }"


def _seeds(n, s=0):
    return np.random.default_rng(s).integers(0, 2 ** 64, size=n, dtype=np.uint64)


def _same(sa, sb):
    for k in ("x", "lp", "accepted", "naccept", "grad", "S"):
        if sa[k] is None:
            continue
        assert np.array_equal(sa[k], sb[k], equal_nan=True), k
    assert sa["step"] == sb["step"]


def _samplers(amh, d):
    Sigma = make_spd(d, seed=3, lo=0.05, hi=0.5)
    return {
        "rwmh_full": (amh.RWMH(amh.MvNormal(np.zeros(d), 0.2 * Sigma)), None),
        "rwmh_scalar": (amh.RWMH(amh.MvNormal(np.zeros(d), 0.01 * amh.I)), None),
        "static_nonsym": (amh.MetropolisHastings(amh.StaticProposal(amh.MvNormal(np.full(d, 0.5), 0.5 * amh.I))), None),
        "rw_mean": (amh.MetropolisHastings(amh.RandomWalkProposal(amh.MvNormal(np.full(d, 0.01), 0.02 * amh.I))), None),
        "mixed": (amh.MetropolisHastings([amh.StaticProposal(amh.Normal(0.5, 1.0)) if i % 2 else amh.RandomWalkProposal(amh.Normal(0, 0.1))
                                          for i in range(d)]), None),
        "ram": (amh.RobustAdaptiveMetropolis(), "warmup"),
        "mala": (amh.MALA(lambda g: amh.MvNormal(0.5 * 2e-3 * g, 2e-3 * amh.I)), "init"),
    }


# ------------------------------------------------------------------ CPU: oracle and host logic
@pytest.mark.parametrize("name", ["rwmh_full", "static_nonsym", "rw_mean", "mixed", "ram", "mala"])
def test_oracle_source_target_restating_rosenbrock_equals_the_catalogue_entry(amh, oracle, name):
    d, n = 6, 96
    spl, mode = _samplers(amh, d)[name]
    cat = amh.RosenbrockTarget(d)
    init = np.random.default_rng(1).normal(0.5, 0.3, size=(d, n)) if mode == "init" else None
    states = []
    for th in (oracle.target(cat.kind, d, cat.blob()), oracle.target_source(d, ROSENBROCK_SRC, cat.blob(), has_gradient=True)):
        run = oracle.run(th, spl.lower(oracle, d), n, _seeds(n, 5), init)
        run.steps(30, warmup=(mode == "warmup"))
        states.append(run.state())
    _same(*states)


def test_oracle_source_stretch_equals_the_catalogue_entry(amh, oracle):
    d, nw, ne = 4, 64, 3
    spl = amh.Ensemble(nw, amh.StretchProposal(amh.MvNormal(np.zeros(d), amh.I)))
    cat = amh.RosenbrockTarget(d)
    states = []
    for th in (oracle.target(cat.kind, d, cat.blob()), oracle.target_source(d, ROSENBROCK_SRC, cat.blob())):
        run = oracle.run(th, spl.lower(oracle, d), nw * ne, _seeds(ne, 6))
        run.steps(25)
        states.append(run.state())
    _same(*states)


def test_oracle_source_iid_normal_equals_the_readme_model(amh, oracle):
    data = np.random.default_rng(1234).normal(0, 1, 30)
    cat = amh.IIDNormalTarget(data)
    n = 64
    init = np.tile(np.array([[0.0], [1.0]]), (1, n))
    states = []
    for th in (oracle.target(cat.kind, 2, cat.blob()), oracle.target_source(2, IID_NORMAL_SRC, data)):
        run = oracle.run(th, amh.RWMH(2).lower(oracle, 2), n, _seeds(n, 7), init)
        run.steps(50)
        states.append(run.state())
    _same(*states)


def test_source_that_does_not_compile_is_an_argument_error_with_the_compiler_log(amh, oracle):
    with pytest.raises(amh.AMHArgumentError) as ei:
        oracle.target_source(3, BROKEN_SRC)
    assert "undefined_symbol" in str(ei.value)
    with pytest.raises(amh.AMHArgumentError):
        oracle.target_source(3, "")
    with pytest.raises(amh.AMHArgumentError):
        oracle.target_source(200, ROSENBROCK_SRC, [1.0, 100.0, 20.0])      # dim <= 128


def test_mala_without_a_gradient_fails_like_the_reference(amh, oracle):
    """MALA.jl:44-50: 'The gradient of the log density function is not defined'"""
    d, n = 3, 8
    th = oracle.target_source(d, FUNNEL_SRC, [0.3, -0.2])
    spl = amh.MALA(lambda g: amh.MvNormal(0.5 * 1e-2 * g, 1e-2 * amh.I))
    with pytest.raises(amh.AMHArgumentError) as ei:
        oracle.run(th, spl.lower(oracle, d), n, _seeds(n), np.zeros((d, n)))
    assert "gradient" in str(ei.value).lower()


def test_density_model_accepts_a_source_target_and_rejects_a_python_closure(amh):
    t = amh.SourceTarget(5, ROSENBROCK_SRC, data=[1.0, 100.0, 20.0], gradient=True)
    m = amh.DensityModel(t)
    assert m.logdensity.dimension() == 5 and m.logdensity.has_gradient
    with pytest.raises(ValueError):
        amh.DensityModel(lambda x: -0.5 * float(x @ x))


def test_nvrtc_translation_unit_compiles_for_sm_100a_without_a_gpu():
    """The library's device code + a user source, exactly as amh_rtc.cu assembles it, through NVRTC offline."""
    try:
        nv = ctypes.CDLL("libnvrtc.so.12")
    except OSError:
        pytest.skip("libnvrtc not on the loader path")
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import gen_rtc_source
    src = (gen_rtc_source.build() + ROSENBROCK_SRC).encode()
    for expr in (b"&amhd::init_kernel<amhd::TUser>", b"&amhd::mh_step_kernel<0, amhd::TUser, 64, 4>",
                 b"&amhh::mala_step_kernel<0, amhd::TUser, 32>"):
        prog = ctypes.c_void_p()
        assert nv.nvrtcCreateProgram(ctypes.byref(prog), src, b"amh_user_target.cu", 0, None, None) == 0
        assert nv.nvrtcAddNameExpression(prog, expr) == 0
        opts = [b"--gpu-architecture=sm_100a", b"--std=c++17", b"--fmad=false"]
        rc = nv.nvrtcCompileProgram(prog, len(opts), (ctypes.c_char_p * len(opts))(*opts))
        n = ctypes.c_size_t()
        nv.nvrtcGetProgramLogSize(prog, ctypes.byref(n))
        log = ctypes.create_string_buffer(max(1, n.value))
        nv.nvrtcGetProgramLog(prog, log)
        assert rc == 0, log.value.decode()[:2000]
        size = ctypes.c_size_t()
        assert nv.nvrtcGetCUBINSize(prog, ctypes.byref(size)) == 0 and size.value > 0
        nv.nvrtcDestroyProgram(ctypes.byref(prog))


# ------------------------------------------------------------------ GPU: NVRTC-compiled targets against the oracle
def _pair_src(cuda, oracle, d, src, data, grad, spl, n, seeds, init=None):
    runs = []
    for eng in (cuda, oracle):
        th = eng.target_source(d, src, data, has_gradient=grad)
        runs.append(eng.run(th, spl.lower(eng, d), n, seeds, init))
    return runs


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["rwmh_full", "rwmh_scalar", "static_nonsym", "rw_mean", "mixed", "ram", "mala"])
def test_gpu_source_target_bit_exact_every_sampler(amh, cuda, oracle, name):
    d, n = 6, 700
    spl, mode = _samplers(amh, d)[name]
    init = np.random.default_rng(1).normal(0.5, 0.3, size=(d, n)) if mode == "init" else None
    rg, ro = _pair_src(cuda, oracle, d, ROSENBROCK_SRC, [1.0, 100.0, 20.0], True, spl, n, _seeds(n, 5), init)
    _same(rg.state(), ro.state())
    for k, per in ((1, 1), (9, 4), (40, 0)):
        rg.steps(k, warmup=(mode == "warmup"), steps_per_launch=per)
        ro.steps(k, warmup=(mode == "warmup"))
        _same(rg.state(), ro.state())
    assert rg.state()["naccept"].sum() > 0


@pytest.mark.gpu
@pytest.mark.parametrize("nw,ne", [(64, 5), (1000, 2), (4096, 1)])
def test_gpu_source_target_stretch_bit_exact(amh, cuda, oracle, nw, ne):
    d = 5
    spl = amh.Ensemble(nw, amh.StretchProposal(amh.MvNormal(np.zeros(d), amh.I)))
    rg, ro = _pair_src(cuda, oracle, d, ROSENBROCK_SRC, [1.0, 100.0, 20.0], False, spl, nw * ne, _seeds(ne, 8))
    _same(rg.state(), ro.state())
    for k in (1, 17):
        rg.steps(k); ro.steps(k)
        _same(rg.state(), ro.state())


@pytest.mark.gpu
def test_gpu_source_target_equals_the_catalogue_kernels(amh, cuda):
    """the NVRTC build of the generic kernel and the ahead-of-time exact-dimension kernel give the same bits"""
    d, n = 10, 2048
    cat = amh.RosenbrockTarget(d)
    spl = amh.RWMH(amh.MvNormal(np.zeros(d), 0.01 * amh.I))
    states = []
    for th in (cuda.target(cat.kind, d, cat.blob()), cuda.target_source(d, ROSENBROCK_SRC, cat.blob())):
        run = cuda.run(th, spl.lower(cuda, d), n, _seeds(n, 9))
        run.steps(100)
        states.append(run.state())
    _same(*states)


@pytest.mark.gpu
def test_gpu_funnel_with_data_and_contract_math_bit_exact(amh, cuda, oracle):
    d, n = 9, 1024
    data = np.random.default_rng(2).normal(size=37)
    spl = amh.RWMH(amh.MvNormal(np.zeros(d), 0.3 * amh.I))
    rg, ro = _pair_src(cuda, oracle, d, FUNNEL_SRC, data, False, spl, n, _seeds(n, 10))
    rg.steps(200); ro.steps(200)
    _same(rg.state(), ro.state())
    acc = rg.state()["naccept"].sum() / (n * 200)
    assert 0.05 < acc < 0.95


@pytest.mark.gpu
def test_gpu_errors_compile_log_and_missing_gradient(amh, cuda):
    with pytest.raises(amh.AMHArgumentError) as ei:
        cuda.target_source(3, BROKEN_SRC)
    assert "undefined_symbol" in str(ei.value)
    d, n = 3, 64
    th = cuda.target_source(d, FUNNEL_SRC, [0.3, -0.2])
    spl = amh.MALA(lambda g: amh.MvNormal(0.5 * 1e-2 * g, 1e-2 * amh.I))
    with pytest.raises(amh.AMHArgumentError) as ei:
        cuda.run(th, spl.lower(cuda, d), n, _seeds(n), np.zeros((d, n)))
    assert "gradient" in str(ei.value).lower()


@pytest.mark.gpu
def test_gpu_sample_api_with_a_source_model_readme_example(amh, cuda, oracle):
    """README.md:26-47 with the model stated as source text: 4 chains, posterior of (mu, sigma) of 30 N(0,1) draws."""
    data = np.random.default_rng(1234).normal(0, 1, 30)
    model = amh.DensityModel(amh.SourceTarget(2, IID_NORMAL_SRC, data))
    spl = amh.RWMH(amh.MvNormal(np.zeros(2), 0.25 * amh.I))
    init = np.tile(np.array([[0.0], [1.0]]), (1, 64))
    ch = amh.sample(model, spl, amh.MCMCB200(device=0), 2000, 64, initial_params=init, discard_initial=500, seed=3,
                    chain_type=amh.Chains)
    ref = amh.sample(model, spl, amh.MCMCB200(device=0), 2000, 64, initial_params=init, discard_initial=500, seed=3,
                     chain_type=amh.Chains, engine=oracle)
    assert np.array_equal(ch.value, ref.value)
    post = ch.value[:, :2, :]
    assert abs(post[:, 0, :].mean() - data.mean()) < 0.1
    assert abs(post[:, 1, :].mean() - data.std()) < 0.15
