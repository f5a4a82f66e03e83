# ref_mcmcthreads.jl -- the reference's own multi-chain CPU path, timed: `sample(model, spl, MCMCThreads(), N, nchains)`
# (README.md:141-147, test/runtests.jl:104-109) on BASELINE config 2 (RWMH, MvNormal d = 32, full-Cholesky proposal), the
# workload bench.py times on the B200.  Prints ONE JSON line in bench.py's `--impl reference` format.
#
#   julia -t auto --project=julia/AdvancedMHB200 bench/ref_mcmcthreads.jl [nchains = 4096] [N = 2000]
#
# STATUS: never executed here -- neither the build container nor the GPU box has a Julia toolchain (BASELINE.md 4), which is
# why bench.py's reference arm times the C++ restatement (oracle) instead and labels it kind = "port".  With Julia
# available this script gives the kind = "reference" number for the same metric: chain-steps/s summed over all chains.
using AdvancedMH, AbstractMCMC, Distributions, LinearAlgebra, LogDensityProblems, MCMCChains, Random, Printf
using AdvancedMHB200        # only for MvNormalTarget (a LogDensityProblems object); no GPU call is made here

nchains = length(ARGS) >= 1 ? parse(Int, ARGS[1]) : 4096
N = length(ARGS) >= 2 ? parse(Int, ARGS[2]) : 2000
d = 32

# the same synthetic problem as bench.py:make_problem (Sigma = Q diag(1..100, log-spaced) Q'); the exact Q differs from
# numpy's (different RNG) -- chain-steps/s does not depend on it
Random.seed!(32)
Q = Matrix(qr(randn(d, d)).Q)
Sigma = Symmetric(Q * Diagonal(exp.(range(0, log(100); length=d))) * Q')
model = AbstractMCMC.LogDensityModel(AdvancedMHB200.MvNormalTarget(zeros(d), Matrix(Sigma)))
spl = RWMH(MvNormal(zeros(d), (2.38^2 / d) * Sigma))
L = cholesky(Sigma).L
init = [L * randn(d) for _ in 1:nchains]

sample(model, spl, MCMCThreads(), 50, min(nchains, 4 * Threads.nthreads()); initial_params=init[1:min(nchains, 4 * Threads.nthreads())],
       chain_type=Any, progress=false)                                   # compile
best = Inf
for rep in 1:3
    t = @elapsed sample(model, spl, MCMCThreads(), N, nchains; initial_params=init, chain_type=Any, progress=false)
    global best = min(best, t)
end
value = nchains * (N - 1) / best                                          # N samples = N - 1 stateful steps per chain
@printf("{\"impl\": \"reference\", \"metric\": \"chain-steps/sec (all chains) on d=32 MvNormal\", \"value\": %.6g, \"unit\": \"chain-steps/s\", \"higher_is_better\": true, \"dtype\": \"f64\", \"data\": \"synthetic\", \"cpu_baseline\": {\"value\": %.6g, \"unit\": \"chain-steps/s\", \"cores\": %d, \"kind\": \"reference\", \"sample\": \"AdvancedMH.jl sample(model, RWMH(MvNormal), MCMCThreads(), %d, %d), best of 3\"}}\n",
        value, value, Threads.nthreads(), N, nchains)
