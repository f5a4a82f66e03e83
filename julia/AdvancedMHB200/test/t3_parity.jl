# t3_parity.jl -- parity tier T3 (SURVEY.md 8c): the UNMODIFIED AdvancedMH.jl on the CPU, consuming the contract stream
# through PhiloxRNG, against the chains the B200 engine produced from the same seeds.
#
#   julia --project=julia/AdvancedMHB200 julia/AdvancedMHB200/test/t3_parity.jl
#   env: AMH_B200_LIB = path of libamh_b200.so, AMH_ORACLE_LIB = path of oracle/libamh_oracle.so
#
# STATUS: never executed (no Julia in the build container or on the GPU box) -- shipped so that a maintainer with Julia
# and a B200 can run it unmodified.  Expected outcome: identical accept / reject decisions for every chain and step and
# states equal to ~1e-12 relative: the normals, exponentials and uniforms are bit-identical (same contract arithmetic),
# while `rand(MvNormal)`'s L*z (BLAS trmv), `logpdf` and Base's `log` / `exp` inside the target differ from the contract's
# operation order in the last bits, so a decision can only flip when |log(alpha) + e| is within rounding of zero.
using AdvancedMH, AbstractMCMC, Distributions, LinearAlgebra, LogDensityProblems, Random, Test
using AdvancedMHB200
include("PhiloxRNG.jl")
using .PhiloxRNGs

const AMHB = AdvancedMHB200

"one reference chain per seed, driven step by step exactly like AbstractMCMC.mcmcsample does (mh-core.jl:76-117)"
function reference_chain(model, spl, seed, d, N; initial_params=nothing, warmup=0)
    rng = PhiloxRNG(seed, d)
    kw = initial_params === nothing ? (;) : (; initial_params)
    t, state = warmup > 0 ? AbstractMCMC.step_warmup(rng, model, spl; kw...) : AbstractMCMC.step(rng, model, spl; kw...)
    xs = [copy(AdvancedMH.getparams(model, state))]; acc = Bool[false]
    for k in 1:(N - 1)
        seekstep!(rng, k)                                              # step k owns blocks [kB, (k+1)B)
        t, state = k <= warmup ? AbstractMCMC.step_warmup(rng, model, spl, state) : AbstractMCMC.step(rng, model, spl, state)
        push!(xs, copy(AdvancedMH.getparams(model, state)))
        push!(acc, t isa AdvancedMH.Transition ? t.accepted : true)
    end
    reduce(hcat, xs), acc
end

"the same chains on the GPU: sample(..., MCMCB200(), N, nchains) with the seeds handed over verbatim"
struct FixedSeeds <: Random.AbstractRNG; seeds::Vector{UInt64}; end
Random.rand(r::FixedSeeds, ::Type{UInt64}, n::Integer) = (@assert n == length(r.seeds); copy(r.seeds))

function compare(name, model, spl, d, N, seeds; initial_params=nothing, warmup=0, rtol=1e-10)
    gpu = sample(FixedSeeds(seeds), model, spl, MCMCB200(), N, length(seeds); initial_params=initial_params === nothing ? nothing : fill(initial_params, length(seeds)),
                 num_warmup=warmup, discard_initial=0, chain_type=Any, progress=false)
    flips = 0
    @testset "$name" begin
        for (c, seed) in enumerate(seeds)
            xs, acc = reference_chain(model, spl, seed, d, N; initial_params, warmup)
            for i in 1:N
                g = gpu[c][i]
                flips += (i > 1 && g.accepted != acc[i])
                g.accepted == acc[i] && @test isapprox(g.params, xs[:, i]; rtol)
            end
        end
        @test flips == 0
    end
end

Random.seed!(1234)
seeds = rand(UInt64, 16)

# RWMH on the reference's own test target (test/runtests.jl:23-31): theta = (mu, sigma), 300 data points
data = randn(300)
iid = AMHB.IIDNormalTarget(data)
compare("RWMH iid-Normal (runtests.jl:76-94)", AbstractMCMC.LogDensityModel(iid), RWMH(MvNormal(zeros(2), I)), 2, 200, seeds)
compare("StaticMH iid-Normal (runtests.jl:56-74)", AbstractMCMC.LogDensityModel(iid), StaticMH(MvNormal([0.0, 1.0], I)), 2, 200, seeds)

# BASELINE config 2 shape: d = 32 full covariance
d = 32
Q = Matrix(qr(randn(d, d)).Q); Sigma = Symmetric(Q * Diagonal(exp.(range(0, log(100); length=d))) * Q')
mvn = AMHB.MvNormalTarget(zeros(d), Matrix(Sigma))
compare("RWMH MvNormal d=32 (config 2)", AbstractMCMC.LogDensityModel(mvn), RWMH(MvNormal(zeros(d), (2.38^2 / d) * Sigma)), d, 100, seeds)

# MALA, issue-95 Gaussian (test/runtests.jl:334-365)
A = inv([1.5 0.35; 0.35 1.0])
gp = AMHB.GaussianPrecisionTarget(A)
s2 = 0.5
compare("MALA Gaussian (runtests.jl:334-365)", AbstractMCMC.LogDensityModel(gp), MALA(g -> MvNormal((s2 / 2) .* g, s2 * I)), 2, 200, seeds;
        initial_params=ones(2))

# RAM warm-up: Givens up/down-dates (RobustAdaptiveMetropolis.jl:153-173), doctest target (:21-36)
ram = AMHB.MvNormalTarget(zeros(2), [1.0 0.5; 0.5 1.0])
compare("RAM warm-up (doctest RAM :17-70)", AbstractMCMC.LogDensityModel(ram), RobustAdaptiveMetropolis(), 2, 300, seeds; initial_params=zeros(2), warmup=300)
