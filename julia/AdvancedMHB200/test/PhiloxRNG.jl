# PhiloxRNG.jl -- TEST INFRASTRUCTURE (parity tier T3, SURVEY.md 8c): an AbstractRNG that makes the UNMODIFIED AdvancedMH.jl
# consume the numerical contract's counter stream (include/amh_contract.h), so that a stock CPU run
# `sample(PhiloxRNG(...), model, sampler, N)` can be compared step by step with the chain the B200 engine produced from
# the same seed.  Every number comes from the CPU oracle's probes (oracle/libamh_oracle.so: amho_probe_philox,
# amho_probe_normal_pair, amho_probe_exponential, amho_probe_u01), i.e. from the contract's own Philox / Box-Muller /
# -log(u) arithmetic -- never from Base's ziggurat.
#
# STATUS: never executed (no Julia in the build container or on the GPU box).  The call sites it serves:
#   randn   src/proposal.jl:25-28 (rand(rng, MvNormal / Normal)), RobustAdaptiveMetropolis.jl:135,193
#   randexp src/mh-core.jl:108, src/MALA.jl:86, src/emcee.jl:93, RobustAdaptiveMetropolis.jl:148
#   rand    src/emcee.jl:81 (uniform), src/emcee.jl:48,52 (Random.Sampler(rng, 1:n-1))
#
# Stream layout (amh_contract.h "stream word budget"), per chain seeded `seed`:
#   MH / MALA / RAM, dimension d.  Contract v1: B = cld(d,2)+1 Philox4x32-10 blocks per step; step k (k = 0 is the initial
#     draw) owns blocks [kB, (k+1)B): block j -> normals z[2j], z[2j+1]; block cld(d,2), word 0 -> the step's exponential.
#     Contract v2 (the library default): B = cld(d,4)+1 Philox4x32-7 blocks; block j -> normals z[4j..4j+3] (two Box-Muller
#     pairs from its 32-bit words); block cld(d,4), 64-bit word 0 -> the exponential.
#     The reference draws d normals then ONE exponential per step and only normals at initialisation, so a sequential
#     consumer advances to the next step after `randexp`, or when a (d+1)-th normal is requested (initial draw -> step 1).
#   stretch move, per ENSEMBLE: initial draw of walker w from stream 1, blocks w*cld(d,2) + j; move i of sweep k owns
#     blocks 2(k*nw + i), 2(k*nw + i) + 1 of stream 0: block 0 word 0 -> partner index, word 1 -> uniform for z,
#     block 1 word 0 -> exponential.
module PhiloxRNGs

using Random

const liboracle = get(ENV, "AMH_ORACLE_LIB", "libamh_oracle")

philox(blk::UInt64, stream::UInt32, seed::UInt64) = begin
    out = Vector{UInt32}(undef, 4)
    ccall((:amho_probe_philox, liboracle), Cvoid, (UInt32, UInt32, UInt32, UInt32, UInt32, UInt32, Ptr{UInt32}),
          blk % UInt32, (blk >> 32) % UInt32, stream, UInt32(0), seed % UInt32, (seed >> 32) % UInt32, out)
    out
end
philox7(blk::UInt64, stream::UInt32, seed::UInt64) = begin          # the step-noise blocks of contract v2
    out = Vector{UInt32}(undef, 4)
    ccall((:amho_probe_philox7, liboracle), Cvoid, (UInt32, UInt32, UInt32, UInt32, UInt32, UInt32, Ptr{UInt32}),
          blk % UInt32, (blk >> 32) % UInt32, stream, UInt32(0), seed % UInt32, (seed >> 32) % UInt32, out)
    out
end
function normal_pair32(wr::UInt32, wa::UInt32)
    a, b = Ref(wr), Ref(wa); z0, z1 = Ref(0.0), Ref(0.0)
    ccall((:amho_probe_normal_pair32, liboracle), Cvoid, (Ptr{UInt32}, Ptr{UInt32}, Ptr{Float64}, Ptr{Float64}, Int64), a, b, z0, z1, 1)
    z0[], z1[]
end
word(b, i) = UInt64(b[2i + 1]) | (UInt64(b[2i + 2]) << 32)          # 64-bit word i (0 or 1) of a block

function normal_pair(b)
    w0, w1 = Ref(word(b, 0)), Ref(word(b, 1)); z0, z1 = Ref(0.0), Ref(0.0)
    ccall((:amho_probe_normal_pair, liboracle), Cvoid, (Ptr{UInt64}, Ptr{UInt64}, Ptr{Float64}, Ptr{Float64}, Int64), w0, w1, z0, z1, 1)
    z0[], z1[]
end
function exponential(w::UInt64)
    x, y = Ref(w), Ref(0.0)
    ccall((:amho_probe_exponential, liboracle), Cvoid, (Ptr{UInt64}, Ptr{Float64}, Int64), x, y, 1)
    y[]
end
function u01(w::UInt64)
    x, y = Ref(w), Ref(0.0)
    ccall((:amho_probe_u01, liboracle), Cvoid, (Ptr{UInt64}, Ptr{Float64}, Int64), x, y, 1)
    y[]
end

"""
    PhiloxRNG(seed, d; contract = 2)         # MetropolisHastings / MALA / RobustAdaptiveMetropolis, dimension d
    PhiloxRNG(seed, d; n_walkers = nw)       # Ensemble(nw, StretchProposal(...)); `seed` is the ENSEMBLE's seed
"""
mutable struct PhiloxRNG <: Random.AbstractRNG
    seed::UInt64
    d::Int
    cv::Int            # contract version of the step noise (1 or 2)
    nw::Int            # 0: chain layout, > 0: ensemble layout
    step::UInt64       # chain layout: current step k; ensemble layout: moves consumed so far (k*nw + i), after the initial draws
    pos::Int           # chain layout: normals consumed in this step; ensemble layout: normals consumed in the initial draw
    init_done::Bool    # ensemble layout: the nw initial draws are over
    cache::Float64     # second normal of the current Box-Muller pair
end
PhiloxRNG(seed::Integer, d::Integer; n_walkers::Integer=0, contract::Integer=2) =
    PhiloxRNG(UInt64(seed), Int(d), Int(contract), Int(n_walkers), 0, 0, false, NaN)

"re-position at the start of step k (chain layout): the hook a step-by-step comparison uses"
seekstep!(r::PhiloxRNG, k::Integer) = (r.step = UInt64(k); r.pos = 0; r)

normal_blocks(r::PhiloxRNG) = r.cv == 2 ? cld(r.d, 4) : cld(r.d, 2)
blocks_per_step(r::PhiloxRNG) = UInt64(normal_blocks(r) + 1)
step_block(r::PhiloxRNG, blk::UInt64) = r.cv == 2 ? philox7(blk, UInt32(0), r.seed) : philox(blk, UInt32(0), r.seed)

function Random.randn(r::PhiloxRNG, ::Type{Float64}=Float64)
    if r.nw == 0
        if r.pos == r.d                         # a (d+1)-th normal: the initial draw is over, this is step + 1
            r.step += 1; r.pos = 0
        end
        if iseven(r.pos)
            if r.cv == 2                        # pair (pos % 4) / 2 of block pos / 4: words (0,1) or (2,3)
                b = step_block(r, r.step * blocks_per_step(r) + UInt64(r.pos ÷ 4))
                h = (r.pos % 4) ÷ 2
                z0, z1 = normal_pair32(b[2h + 1], b[2h + 2])
            else
                z0, z1 = normal_pair(step_block(r, r.step * blocks_per_step(r) + UInt64(r.pos ÷ 2)))
            end
            r.cache = z1; r.pos += 1
            return z0
        end
        r.pos += 1
        return r.cache
    else                                        # initial draw of walker w = pos ÷ d (emcee.jl:29-34), stream 1
        w, p = divrem(r.pos, r.d)
        if iseven(p)
            z0, z1 = normal_pair(philox(UInt64(w) * UInt64(cld(r.d, 2)) + UInt64(p ÷ 2), UInt32(1), r.seed))
            r.cache = z1; r.pos += 1
            return z0
        end
        r.pos += 1
        return r.cache
    end
end

function Random.randexp(r::PhiloxRNG, ::Type{Float64}=Float64)
    if r.nw == 0
        r.pos == r.d || r.step == 0 && r.pos == 0 ||
            error("PhiloxRNG: randexp after $(r.pos) of $(r.d) normals -- not the reference's per-step consumption")
        e = exponential(word(step_block(r, r.step * blocks_per_step(r) + UInt64(normal_blocks(r))), 0))
        r.step += 1; r.pos = 0                  # the exponential closes the step (mh-core.jl:108)
        return e
    else
        e = exponential(word(philox(2 * r.step + 1, UInt32(0), r.seed), 0))
        r.step += 1                             # the exponential closes the move (emcee.jl:93)
        return e
    end
end

# rand(rng) :: Float64 in (0, 1)  (emcee.jl:81: the stretch factor's uniform = word 1 of the move's first block)
function Random.rand(r::PhiloxRNG, ::Random.SamplerTrivial{Random.CloseOpen01{Float64}})
    r.nw > 0 || error("PhiloxRNG: the chain layout has no uniform slot (only the stretch move draws one)")
    u01(word(philox(2 * r.step, UInt32(0), r.seed), 1))
end

# rand(rng, Random.Sampler(rng, 1:(n-1)))  (emcee.jl:48,52): floor(w * n / 2^64) + 1, the contract's `bounded`
function Random.rand(r::PhiloxRNG, sp::Random.SamplerRangeNDL{UInt64,Int})
    r.nw > 0 || error("PhiloxRNG: the chain layout has no integer slot")
    r.init_done = true
    w = word(philox(2 * r.step, UInt32(0), r.seed), 0)
    n = UInt64(sp.s)                            # length of the range
    Int(((UInt128(w) * UInt128(n)) >> 64) % UInt64) + first(sp.a)
end
Random.rand(r::PhiloxRNG, sp::Random.SamplerRangeNDL) = rand(r, Random.SamplerRangeNDL{UInt64,Int}(Int(first(sp.a)), UInt64(sp.s)))

# raw words for anything else (e.g. `rand(rng, UInt, nchains)`, the per-chain seeding of AbstractMCMC): not part of the
# contract stream -- served from stream 255 so that it can never collide with a slot above
Random.rng_native_52(::PhiloxRNG) = UInt64
function Random.rand(r::PhiloxRNG, ::Random.SamplerType{UInt64})
    r.step += 1
    word(philox(r.step, UInt32(255), r.seed), 0)
end

export PhiloxRNG, seekstep!
end # module
