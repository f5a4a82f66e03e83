# AdvancedMHB200.jl -- thin Julia shim over libamh_b200.so (include/amh.h).
#
# STATUS: written against the C ABI and the AbstractMCMC >= 5.6 interface, but NEVER EXECUTED: there is no Julia
# toolchain in the build container nor on the GPU box.  The tested equivalent of every call below is the Python
# ctypes mirror (advancedmh.jl_b200/_capi.py, sampling.py).  Treat this file as the reference-side binding a
# maintainer would review, not as verified code.
#
# What it adds to an unmodified AdvancedMH.jl:
#   * `MCMCB200 <: AbstractMCMC.AbstractMCMCEnsemble`: sample(model, sampler, MCMCB200(ngpus = 8), N, nchains; kw...)
#     -- one process, N GPUs, through the library's multi-GPU job (amh_job_*, include/amh.h)
#   * catalogue targets (`MvNormalTarget`, ...) that ALSO implement LogDensityProblems, so the same object runs
#     through stock MCMCThreads() for CPU comparison (bench/ref_mcmcthreads.jl)
# NOT in this module: `PhiloxRNG`, the AbstractRNG that makes stock AdvancedMH consume the contract stream for the T3
# parity run.  It calls the CPU oracle's probes, i.e. it is test infrastructure, and lives in test/PhiloxRNG.jl.
module AdvancedMHB200

using AbstractMCMC, AdvancedMH, Distributions, LinearAlgebra, LogDensityProblems, Random

const libamh = get(ENV, "AMH_B200_LIB", "libamh_b200")

# ---------------------------------------------------------------- ABI constants (include/amh.h)
const AMH_OK = Int32(0)
const TARGET_IID_NORMAL, TARGET_MVNORMAL, TARGET_ROSENBROCK, TARGET_LOGISTIC, TARGET_GAUSS_PREC, TARGET_NIG_TOY, TARGET_NIG_TOY_LOG = Int32.(1:7)
const TARGET_USER = Int32(100)
const SAMPLER_STATIC, SAMPLER_RW, SAMPLER_STRETCH, SAMPLER_MALA, SAMPLER_RAM, SAMPLER_MIXED = Int32.(1:6)
const COV_SCALAR, COV_DIAG, COV_FULL, COV_COMPONENTS = Int32.(1:4)
const FAM_NORMAL, FAM_INVGAMMA, FAM_GAMMA, FAM_UNIFORM, FAM_EXPONENTIAL, FAM_LOGNORMAL = Int32.(1:6)

struct Component              # struct amh_component: one univariate law of an array proposal (proposal.jl:26-35)
    family::Int32; rw::Int32; symmetric::Int32; reserved::Int32
    p0::Float64; p1::Float64; logc::Float64
end

struct SamplerDesc            # struct amh_sampler_desc, field for field
    kind::Int32; dim::Int32; symmetric::Int32; cov_kind::Int32
    mean::Ptr{Float64}; scale::Ptr{Float64}
    stretch_a::Float64; n_walkers::Int64
    mala_sigma2::Float64; mala_drift::Float64
    ram_alpha::Float64; ram_gamma::Float64; ram_eig_lo::Float64; ram_eig_hi::Float64
    ram_S0::Ptr{Float64}
    components::Ptr{Component}
    contract::Int32; precision::Int32     # precision: 0 = fp64 (bit-exact), 1 = split-bf16 tensor path (MALA x logistic, opt-in)
end
"version of the numerical contract new runs are created under: 0 = the library default (v2), 1 = v1, 2 = v2"
const CONTRACT = Ref(Int32(0))
"arithmetic of the design-matrix contractions: 0 = fp64 (bit-exact, default), 1 = split-bf16 tensor path (opt-in; MALA x logistic x dim 128)"
const PRECISION = Ref(Int32(0))
SamplerDesc(kind, dim, symmetric, cov_kind, mean, scale, stretch_a, n_walkers, mala_sigma2, mala_drift, ram_alpha, ram_gamma,
            ram_eig_lo, ram_eig_hi, ram_S0, components) =
    SamplerDesc(kind, dim, symmetric, cov_kind, mean, scale, stretch_a, n_walkers, mala_sigma2, mala_drift, ram_alpha, ram_gamma,
                ram_eig_lo, ram_eig_hi, ram_S0, components, CONTRACT[], PRECISION[])

struct Summary                # struct amh_summary
    n_saved::Int64; n_steps::Int64; accept_rate::Float64
    mean::Ptr{Float64}; var::Ptr{Float64}; chain_mean::Ptr{Float64}
end

function check(rc::Int32)
    rc == AMH_OK && return nothing
    msg = unsafe_string(ccall((:amh_last_error, libamh), Cstring, ()))
    rc == 1 || rc == 3 ? throw(ArgumentError(msg)) : error(msg)      # AMH_ERR_INVALID/UNSUPPORTED -> ArgumentError
end

# ---------------------------------------------------------------- device-target catalogue
abstract type DeviceTarget end

"""logpdf(MvNormal(mu, Sigma), x); blob = [c0, mu, U packed by rows], U'U = inv(Sigma)"""
struct MvNormalTarget <: DeviceTarget
    mu::Vector{Float64}
    Sigma::Matrix{Float64}
end
kind(::MvNormalTarget) = TARGET_MVNORMAL
function blob(t::MvNormalTarget)
    d = length(t.mu)
    C = cholesky(Symmetric(t.Sigma)).L
    U = inv(C)                                   # lower triangular
    c0 = -0.5 * (d * log(2pi) + 2sum(log, diag(C)))
    vcat(c0, t.mu, [U[i, j] for i in 1:d for j in 1:i])
end
LogDensityProblems.logdensity(t::MvNormalTarget, x) = logpdf(MvNormal(t.mu, t.Sigma), x)
LogDensityProblems.dimension(t::MvNormalTarget) = length(t.mu)
LogDensityProblems.capabilities(::Type{MvNormalTarget}) = LogDensityProblems.LogDensityOrder{1}()
LogDensityProblems.logdensity_and_gradient(t::MvNormalTarget, x) =
    (LogDensityProblems.logdensity(t, x), -(t.Sigma \ (x - t.mu)))

"""-x'Ax/2 (test/runtests.jl:335-347 `TheNormalLogDensity`)"""
struct GaussianPrecisionTarget <: DeviceTarget
    A::Matrix{Float64}
end
kind(::GaussianPrecisionTarget) = TARGET_GAUSS_PREC
blob(t::GaussianPrecisionTarget) = vec(permutedims(t.A))        # row-major
LogDensityProblems.logdensity(t::GaussianPrecisionTarget, x) = -dot(x, t.A, x) / 2
LogDensityProblems.dimension(t::GaussianPrecisionTarget) = size(t.A, 1)
LogDensityProblems.capabilities(::Type{GaussianPrecisionTarget}) = LogDensityProblems.LogDensityOrder{1}()
LogDensityProblems.logdensity_and_gradient(t::GaussianPrecisionTarget, x) = (-dot(x, t.A, x) / 2, -t.A * x)

"""sum(logpdf.(Normal(mu, sigma), data)) on sigma >= 0 (README.md:26-31)"""
struct IIDNormalTarget <: DeviceTarget
    data::Vector{Float64}
end
kind(::IIDNormalTarget) = TARGET_IID_NORMAL
blob(t::IIDNormalTarget) = t.data
LogDensityProblems.logdensity(t::IIDNormalTarget, th) = th[2] >= 0 ? sum(logpdf.(Normal(th[1], th[2]), t.data)) : -Inf
LogDensityProblems.dimension(::IIDNormalTarget) = 2
LogDensityProblems.capabilities(::Type{IIDNormalTarget}) = LogDensityProblems.LogDensityOrder{0}()

struct RosenbrockTarget <: DeviceTarget
    dim::Int; a::Float64; b::Float64; s::Float64
end
RosenbrockTarget(d) = RosenbrockTarget(d, 1.0, 100.0, 20.0)
kind(::RosenbrockTarget) = TARGET_ROSENBROCK
blob(t::RosenbrockTarget) = [t.a, t.b, t.s]
LogDensityProblems.logdensity(t::RosenbrockTarget, x) =
    -sum(t.b * (x[i + 1] - x[i]^2)^2 + (t.a - x[i])^2 for i in 1:(t.dim - 1)) / t.s
LogDensityProblems.dimension(t::RosenbrockTarget) = t.dim
LogDensityProblems.capabilities(::Type{RosenbrockTarget}) = LogDensityProblems.LogDensityOrder{0}()

"""Bayesian logistic regression, prior N(0, tau^2 I): sum_i [y_i eta_i - log1pexp(eta_i)] - |beta|^2 / (2 tau^2), eta = X beta
(BASELINE config 4); blob = [tau, X row-major, y]"""
struct LogisticRegressionTarget <: DeviceTarget
    X::Matrix{Float64}      # n x d
    y::Vector{Float64}
    tau::Float64
end
LogisticRegressionTarget(X, y; tau=10.0) = LogisticRegressionTarget(Matrix{Float64}(X), collect(Float64, y), Float64(tau))
kind(::LogisticRegressionTarget) = TARGET_LOGISTIC
blob(t::LogisticRegressionTarget) = vcat(t.tau, vec(permutedims(t.X)), t.y)
log1pexp(x) = x > 0 ? x + log1p(exp(-x)) : log1p(exp(x))
function LogDensityProblems.logdensity(t::LogisticRegressionTarget, b)
    eta = t.X * b
    sum(t.y .* eta .- log1pexp.(eta)) - dot(b, b) / (2 * t.tau^2)
end
LogDensityProblems.dimension(t::LogisticRegressionTarget) = size(t.X, 2)
LogDensityProblems.capabilities(::Type{LogisticRegressionTarget}) = LogDensityProblems.LogDensityOrder{1}()
function LogDensityProblems.logdensity_and_gradient(t::LogisticRegressionTarget, b)
    eta = t.X * b
    (sum(t.y .* eta .- log1pexp.(eta)) - dot(b, b) / (2 * t.tau^2), t.X' * (t.y .- 1 ./ (1 .+ exp.(-eta))) .- b ./ t.tau^2)
end

"""The emcee example of the reference's tests: s ~ InverseGamma(alpha, beta), m ~ N(0, s), y_i ~ N(m, s)
(test/emcee.jl:5-15), or in (log s, m) with the Jacobian term (`log_space = true`, test/emcee.jl:46-56)"""
struct NormalInverseGammaToy <: DeviceTarget
    obs::Vector{Float64}
    alpha::Float64
    beta::Float64
    log_space::Bool
end
NormalInverseGammaToy(obs=[1.5, 2.0]; alpha=2.0, beta=3.0, log_space=false) =
    NormalInverseGammaToy(collect(Float64, obs), Float64(alpha), Float64(beta), log_space)
kind(t::NormalInverseGammaToy) = t.log_space ? TARGET_NIG_TOY_LOG : TARGET_NIG_TOY
blob(t::NormalInverseGammaToy) = vcat(t.alpha, t.beta, t.alpha * log(t.beta) - first(logabsgamma(t.alpha)), t.obs)
function LogDensityProblems.logdensity(t::NormalInverseGammaToy, th)
    s, m = t.log_space ? (exp(th[1]), th[2]) : (th[1], th[2])
    s > 0 || return -Inf
    lp = logpdf(InverseGamma(t.alpha, t.beta), s) + logpdf(Normal(0, sqrt(s)), m) + sum(logpdf.(Normal(m, sqrt(s)), t.obs))
    t.log_space ? lp + th[1] : lp
end
LogDensityProblems.dimension(::NormalInverseGammaToy) = 2
LogDensityProblems.capabilities(::Type{NormalInverseGammaToy}) = LogDensityProblems.LogDensityOrder{0}()

"""
    SourceTarget(dim, source; data = Float64[], gradient = false, logdensity = nothing)

A log-density stated as C++ source text (include/amh_user_target.h), compiled for the device by NVRTC inside the
library (`amh_target_create_source`): the route from the catalogue to `DensityModel(f)` for an arbitrary `f`
(AdvancedMH.jl src/AdvancedMH.jl:52-54).  `logdensity` may carry the Julia closure the text restates, so that the same
object also runs through stock `MCMCThreads()`.
"""
struct SourceTarget{F} <: DeviceTarget
    dim::Int
    source::String
    data::Vector{Float64}
    gradient::Bool
    logdensity::F
end
SourceTarget(dim, source; data=Float64[], gradient=false, logdensity=nothing) =
    SourceTarget(Int(dim), String(source), collect(Float64, data), gradient, logdensity)
kind(::SourceTarget) = TARGET_USER
blob(t::SourceTarget) = t.data
LogDensityProblems.logdensity(t::SourceTarget, x) =
    t.logdensity === nothing ? throw(ArgumentError("this SourceTarget carries no Julia closure")) : t.logdensity(x)
LogDensityProblems.dimension(t::SourceTarget) = t.dim
LogDensityProblems.capabilities(::Type{<:SourceTarget}) = LogDensityProblems.LogDensityOrder{0}()

# the job's target: catalogue entry (kind, blob) or source text; the library broadcasts it to every device of the job
function create_target(job, target::DeviceTarget, d)
    b = blob(target)
    GC.@preserve b check(ccall((:amh_job_target_create, libamh), Int32, (Ptr{Cvoid}, Int32, Int32, Ptr{Float64}, Int64),
                               job, kind(target), d, b, length(b)))
end
function create_target(job, target::SourceTarget, d)
    b = target.data
    GC.@preserve b check(ccall((:amh_job_target_create_source, libamh), Int32,
                               (Ptr{Cvoid}, Int32, Cstring, Int32, Ptr{Float64}, Int64),
                               job, d, target.source, target.gradient, isempty(b) ? C_NULL : pointer(b), length(b)))
end

# a DensityModel / LogDensityModel must wrap a catalogue target; anything else cannot run on the device
unwrap(m::AdvancedMH.DensityModel) = unwrap(m.logdensity)
unwrap(m::AbstractMCMC.LogDensityModel) = unwrap(m.logdensity)
unwrap(t::DeviceTarget) = t
unwrap(x) = throw(ArgumentError("MCMCB200 needs a catalogue device target, got $(typeof(x)); arbitrary closures cannot run in a CUDA kernel and there is no CPU fallback"))

# ---------------------------------------------------------------- sampler lowering (unchanged AdvancedMH constructors)
function gaussian(p)      # -> (cov_kind, mean or nothing, scale)
    if p isa MvNormal
        S = p.Σ
        mu = all(iszero, mean(p)) ? nothing : collect(Float64, mean(p))
        S isa Distributions.PDMats.ScalMat && return (COV_SCALAR, mu, [sqrt(S.value)])
        S isa Distributions.PDMats.PDiagMat && return (COV_DIAG, mu, sqrt.(S.diag))
        L = cholesky(Symmetric(Matrix(S))).L
        return (COV_FULL, mu, [L[i, j] for i in 1:size(L, 1) for j in 1:i])
    elseif p isa Normal
        return (COV_SCALAR, iszero(p.μ) ? nothing : [p.μ], [p.σ])
    elseif p isa AbstractVector{<:Normal}
        mu = [q.μ for q in p]
        return (COV_DIAG, all(iszero, mu) ? nothing : mu, [q.σ for q in p])
    end
    throw(ArgumentError("unsupported proposal on the device path: $(typeof(p))"))
end

issym(::AdvancedMH.Proposal{S}) where {S} = S       # the `issymmetric` type parameter (proposal.jl:1-21)

struct Lowered
    desc::SamplerDesc
    keep::Vector{Any}       # arrays the desc points into
end

# one univariate law -> (family, p0, p1, logc) in the parametrisation of include/amh_contract.h
component(q::Normal) = (FAM_NORMAL, q.μ, q.σ, -log(q.σ) - log(2π) / 2)
component(q::LogNormal) = (FAM_LOGNORMAL, q.μ, q.σ, -log(q.σ) - log(2π) / 2)
component(q::InverseGamma) = (FAM_INVGAMMA, shape(q), scale(q), shape(q) * log(scale(q)) - first(logabsgamma(shape(q))))
component(q::Gamma) = (FAM_GAMMA, shape(q), scale(q), -shape(q) * log(scale(q)) - first(logabsgamma(shape(q))))
component(q::Uniform) = (FAM_UNIFORM, q.a, q.b, -log(q.b - q.a))
component(q::Exponential) = (FAM_EXPONENTIAL, scale(q), 0.0, -log(scale(q)))
component(q) = throw(ArgumentError("unsupported univariate proposal law on the device path: $(typeof(q))"))
logabsgamma(x) = Distributions.SpecialFunctions.logabsgamma(x)

function components_desc(kind, d, sym, comps::Vector{Component}, stretch_a=2.0, n_walkers=0)
    length(comps) == d || throw(ArgumentError("proposal dimension $(length(comps)) != model dimension $d"))
    Lowered(SamplerDesc(kind, d, sym, COV_COMPONENTS, C_NULL, C_NULL, stretch_a, n_walkers, 0.0, 0.0,
                        0.234, 0.6, 0.0, Inf, C_NULL, pointer(comps)), Any[comps])
end

function lower(spl::AdvancedMH.MetropolisHastings, d)
    p = spl.proposal
    if p isa Union{AbstractVector{<:AdvancedMH.Proposal},NamedTuple}
        # array / NamedTuple of proposals, one univariate law per coordinate (proposal.jl:132-175, 199-240)
        comps = [Component(component(q.proposal)[1], q isa AdvancedMH.RandomWalkProposal, issym(q), 0,
                           component(q.proposal)[2:4]...) for q in values(p)]
        return components_desc(SAMPLER_MIXED, d, 0, comps)
    end
    p isa Union{AdvancedMH.StaticProposal,AdvancedMH.RandomWalkProposal} ||
        throw(ArgumentError("function-valued proposals are host-only"))
    k = p isa AdvancedMH.RandomWalkProposal ? SAMPLER_RW : SAMPLER_STATIC
    if p.proposal isa AbstractVector{<:UnivariateDistribution} && !(p.proposal isa AbstractVector{<:Normal})
        # StaticProposal([Normal(0,1), InverseGamma(2,3)]) (README.md:106; proposal.jl:26-35)
        comps = [Component(component(q)[1], 0, 0, 0, component(q)[2:4]...) for q in p.proposal]
        return components_desc(k, d, issym(p), comps)
    end
    ck, mu, sc = gaussian(p.proposal)
    keep = Any[mu, sc]
    Lowered(SamplerDesc(k, d, issym(p), ck, mu === nothing ? C_NULL : pointer(mu), pointer(sc), 2.0, 0, 0.0, 0.0,
                        0.234, 0.6, 0.0, Inf, C_NULL, C_NULL), keep)
end

function lower(spl::AdvancedMH.Ensemble, d)
    sp = spl.proposal::AdvancedMH.StretchProposal
    if sp.proposal isa AbstractVector{<:UnivariateDistribution} && !(sp.proposal isa AbstractVector{<:Normal})
        # StretchProposal([InverseGamma(2,3), Normal(0,1)]) (test/emcee.jl:19): the law of the initial draw
        comps = [Component(component(q)[1], 0, 0, 0, component(q)[2:4]...) for q in sp.proposal]
        return components_desc(SAMPLER_STRETCH, d, 0, comps, sp.stretch_length, spl.n_walkers)
    end
    ck, mu, sc = try gaussian(sp.proposal) catch; (COV_SCALAR, nothing, nothing) end
    keep = Any[mu, sc]
    Lowered(SamplerDesc(SAMPLER_STRETCH, d, 0, ck, mu === nothing ? C_NULL : pointer(mu),
                        sc === nothing ? C_NULL : pointer(sc), sp.stretch_length, spl.n_walkers, 0.0, 0.0,
                        0.234, 0.6, 0.0, Inf, C_NULL, C_NULL), keep)
end

function lower(spl::AdvancedMH.MALA, d)
    # recover (sigma2, drift) of g -> MvNormal(drift*g, sigma2*I) by probing the closure on the host: p(0) gives sigma2,
    # p(e_i) the drift coefficient of every coordinate, p(2 e_i) checks linearity; anything else cannot be lowered
    # (same checks as the Python mirror, samplers.py MALA.probe)
    f = spl.proposal.proposal
    bad() = throw(ArgumentError("MALA on the device needs proposal(g) = MvNormal(c*g, sigma2*I)"))
    scal(p) = p isa MvNormal && p.Σ isa Distributions.PDMats.ScalMat && length(p) == d
    p0 = f(zeros(d))
    (scal(p0) && all(iszero, mean(p0))) || bad()
    sigma2 = p0.Σ.value
    c = nothing
    for i in 1:d
        e = zeros(d); e[i] = 1.0
        p1, p2 = f(e), f(2 .* e)
        (scal(p1) && scal(p2) && p1.Σ.value == sigma2) || bad()
        m1, m2 = mean(p1), mean(p2)
        ci = m1[i]
        (all(iszero, m1[1:d .!= i]) && abs(m2[i] - 2ci) <= 1e-12 * max(1.0, abs(ci))) || bad()
        c === nothing ? (c = ci) : (ci == c || bad())
    end
    Lowered(SamplerDesc(SAMPLER_MALA, d, 0, COV_SCALAR, C_NULL, C_NULL, 2.0, 0, sigma2, c,
                        0.234, 0.6, 0.0, Inf, C_NULL, C_NULL), Any[])
end

function lower(spl::AdvancedMH.RobustAdaptiveMetropolis, d)
    S0 = spl.S === nothing ? nothing : vec(permutedims(Matrix{Float64}(spl.S)))
    spl.S === nothing || size(spl.S) == (d, d) || throw(ArgumentError("The provided `S` has the wrong dimensionality."))
    Lowered(SamplerDesc(SAMPLER_RAM, d, 0, COV_SCALAR, C_NULL, C_NULL, 2.0, 0, 0.0, 0.0, spl.α, spl.γ,
                        spl.eigenvalue_lower_bound, spl.eigenvalue_upper_bound, S0 === nothing ? C_NULL : pointer(S0), C_NULL),
            Any[S0])
end

# ---------------------------------------------------------------- the ensemble type
"""
    MCMCB200(; ngpus = 1, devices = nothing, ignore_failed_downdates = false)

Run all chains in lock-step on `ngpus` B200s of this box, from this one process: the drop-in for `MCMCThreads()` /
`MCMCDistributed()` (AdvancedMH.jl src/AdvancedMH.jl:30, README.md:135-148).  The library shards the chains in contiguous
blocks (an `Ensemble` stays on one GPU), broadcasts the target's fixed data once (NCCL over NVLink), steps every block on
its own stream without any per-step collective, and every GPU writes its column block of the result array directly.
`sample(model, sampler, MCMCB200(ngpus = 8), N, nchains; kw...)` keeps AbstractMCMC's keywords: `initial_params` (one
entry per chain), `discard_initial`, `thinning`, `num_warmup`, `chain_type`, `param_names`.
"""
Base.@kwdef struct MCMCB200 <: AbstractMCMC.AbstractMCMCEnsemble
    ngpus::Int = 1
    devices::Union{Nothing,Vector{Int32}} = nothing       # CUDA device indices, default 0:ngpus-1
    ignore_failed_downdates::Bool = false                 # RAM: keep the samples instead of throwing PosDefException
    dtype::Symbol = :fp64                                 # :bf16x2 = opt-in tensor-core arithmetic with a stated tolerance (MALA x logistic)
end

# (n, d) column-major == the device layout X[dim][chain] with chains fastest.  One entry per chain; for an Ensemble an
# entry is the vector of its n_walkers walker positions (emcee.jl:29-34 returns exactly that shape).
function initial_matrix(initial_params, nchains, nw, d)
    initial_params === nothing && return nothing
    length(initial_params) == nchains || throw(ArgumentError("initial_params must have one entry per chain"))
    init = Matrix{Float64}(undef, nchains * nw, d)
    for (c, p) in enumerate(initial_params)
        if nw == 1
            length(p) == d || throw(ArgumentError("initial_params entry has length $(length(p)), model dimension is $d"))
            init[c, :] .= p
        else
            length(p) == nw || throw(ArgumentError("an Ensemble needs n_walkers initial positions per chain"))
            for (w, q) in enumerate(p)
                init[(c - 1) * nw + w, :] .= q
            end
        end
    end
    init
end

function AbstractMCMC.mcmcsample(rng::Random.AbstractRNG, model::AbstractMCMC.AbstractModel,
                                 sampler::AdvancedMH.MHSampler, par::MCMCB200, N::Integer, nchains::Integer;
                                 initial_params=nothing, num_warmup::Integer=0,
                                 discard_initial::Integer=num_warmup, thinning::Integer=1,
                                 chain_type::Type=Any, kwargs...)
    target = unwrap(model)
    d = LogDensityProblems.dimension(target)
    nw = sampler isa AdvancedMH.Ensemble ? sampler.n_walkers : 1
    n = nchains * nw
    seeds = rand(rng, UInt64, nchains)                         # exactly AbstractMCMC's per-chain seeding
    PRECISION[] = par.dtype === :bf16x2 ? Int32(1) : Int32(0)
    low = lower(sampler, d)
    init = initial_matrix(initial_params, nchains, nw, d)
    out = Array{Float64}(undef, n, d + 1, N)                   # C order [N][d+1][n]: every GPU fills its block of columns
    acc = Array{UInt8}(undef, n, N)
    devs = par.devices === nothing ? C_NULL : pointer(par.devices)
    job = Ref{Ptr{Cvoid}}(C_NULL)
    nfailed = Ref{Int64}(0); first_failed = Ref{Int64}(-1)
    GC.@preserve low seeds init out acc par begin
        check(ccall((:amh_job_create, libamh), Int32, (Int32, Ptr{Int32}, Ptr{Ptr{Cvoid}}), par.ngpus, devs, job))
        try
            create_target(job[], target, d)
            check(ccall((:amh_job_sampler_create, libamh), Int32, (Ptr{Cvoid}, Ref{SamplerDesc}), job[], low.desc))
            check(ccall((:amh_job_run_create, libamh), Int32, (Ptr{Cvoid}, Int64, Ptr{UInt64}, Ptr{Float64}, Int64),
                        job[], n, seeds, init === nothing ? C_NULL : pointer(init), 0))
            check(ccall((:amh_job_run_sample, libamh), Int32,
                        (Ptr{Cvoid}, Int64, Int64, Int64, Int64, Ptr{Float64}, Ptr{UInt8}, Ptr{Summary}),
                        job[], N, discard_initial, thinning, num_warmup, out, acc, C_NULL))
            if sampler isa AdvancedMH.RobustAdaptiveMetropolis
                check(ccall((:amh_job_run_ram_failed, libamh), Int32, (Ptr{Cvoid}, Ptr{Int64}, Ptr{Int64}, Ptr{UInt8}),
                            job[], nfailed, first_failed, C_NULL))
            end
        finally
            ccall((:amh_job_destroy, libamh), Int32, (Ptr{Cvoid},), job[])       # releases run, sampler, target, contexts
        end
    end
    # lowrankdowndate throws PosDefException and aborts `sample` (RobustAdaptiveMetropolis.jl:170); so does this path
    if nfailed[] > 0 && !par.ignore_failed_downdates
        throw(LinearAlgebra.PosDefException(Int(first_failed[]) + 1))
    end
    # hand the arrays to the reference's own bundling, then chainsstack exactly like AbstractMCMC does
    # (src/AdvancedMH.jl:80-123, ext/AdvancedMHMCMCChainsExt.jl).  An Ensemble sample is the Vector of ALL its walkers'
    # Transitions (emcee.jl:14-24), which is what the Chains extension bundles (:80-121).
    transition(col, i) = AdvancedMH.Transition(out[col, 1:d, i], out[col, d + 1, i], acc[col, i] != 0)
    chains = map(1:nchains) do c
        ts = nw == 1 ? [transition(c, i) for i in 1:N] :
                       [[transition((c - 1) * nw + w, i) for w in 1:nw] for i in 1:N]
        AbstractMCMC.bundle_samples(ts, model, sampler, nothing, chain_type; discard_initial, thinning, kwargs...)
    end
    return AbstractMCMC.chainsstack(AbstractMCMC.tighten_eltype(chains))
end

export MCMCB200, MvNormalTarget, GaussianPrecisionTarget, IIDNormalTarget, RosenbrockTarget, LogisticRegressionTarget,
       NormalInverseGammaToy, SourceTarget

end # module
