/* amh_launch_mala_logistic.cu -- K3L: MALA step (MALA.jl:54-93) on the Bayesian logistic-regression target with
 * MANY rows (BASELINE config 4: d = 128, 10^4 rows, 16 384 chains): the one place on the hot path where the work
 * is a real dense contraction, so it runs on the FP64 tensor cores.
 *
 * value-and-gradient of the target for 8 chains at a time is two chained GEMMs with a nonlinearity in between
 * (attention-shaped):
 *     eta[rows x 8]  = X[rows x d] . C[d x 8]                       GEMM1, K = d
 *     t = y eta - log1pexp(eta),  r = y - sigmoid(eta)              per (row, chain), contract math
 *     gg[d x 8]      = X'[d x rows] . r[rows x 8]                   GEMM2, K = rows
 * A warp owns 8 chains (one DMMA n-tile) for the whole step and streams X in 8-row blocks: per block 32 DMMAs for
 * GEMM1 (k over d), 2 x (exp, log, 2 divisions) per lane, 32 DMMAs for GEMM2 whose 16 accumulator tiles stay in
 * registers for all 1250 blocks.  DMMA accumulates k as a sequential IEEE fma chain (tools/ubench/dmma_probe.cu), so
 * eta and gg are bit-identical to the oracle's row-by-row loops; the log-likelihood is summed as the contract's
 * 8 interleaved partial sums (one per fragment row) + warp-shuffle tree.
 *
 * X blocks (8 KB, contiguous in the row-major design matrix) are pulled into a 4-stage shared-memory ring by the
 * TMA engine (one cp.async.bulk per padded row, full/empty mbarriers); all warps of the CTA consume the same stage.
 * Candidate / gradient vectors of the chains live in global memory (L2): a step is ~3 ms of FP64 tensor work, the
 * 4 KB of per-chain state traffic is noise.
 */
#include "amh_params.cuh"
#include "amh_fastmath.cuh"

namespace amhh {
using namespace amhd;

struct MalaLArgs {
    ChainState st;
    SaveArgs sv;
    int nsteps;
    unsigned long long step0;
    double sigma, sigma2, drift;
    const double* Xp;          /* [nblk*8][D] row-major, zero padded */
    const double* yp;          /* [nblk*8] */
    long long nrows;
    int nblk;
    int nst;                   /* stages of the X ring (even) */
    double inv2tau2, invtau2;
    double* Xc;                /* [D][pitch] candidate             */
    double* Gc;                /* [D][pitch] gradient at candidate */
    const double* dscale;      /* [D] per-coordinate noise scale: RW: the proposal's standard deviations; MALA: sigma; 0 in the padding */
    /* the run's dimension d <= D (features d..D-1 are padding: zero columns of X, zero scale, zero state) and its noise
     * blocks per step / the exponential's block within a step under the run's contract version */
    int d_real;
    unsigned long long blocks_per_step, exp_block;
    /* RW variant with a FULL-covariance proposal: DMMA A fragments of its lower Cholesky factor, padded to D x D
     * (frag[tile(mb, kb)][lane] = L[8 mb + lane / 4][4 kb + lane % 4], tiles kb = 0 .. 2 mb + 1 of row block mb); else NULL */
    const double* Lf;
};

__device__ __forceinline__ unsigned l_smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void l_mbar_init(unsigned long long* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(l_smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void l_mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(l_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void l_mbar_arrive(unsigned long long* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(l_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void l_mbar_wait(unsigned long long* bar, unsigned phase) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "LWAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra LDONE;\n"
        "bra LWAIT;\n"
        "LDONE:\n"
        "}\n" ::"r"(l_smem_u32(bar)), "r"(phase) : "memory");
}
__device__ __forceinline__ void l_bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(l_smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(l_smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void l_dmma(double& d0, double& d1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

constexpr int kLPB = 8;        /* row pitch of the per-warp candidate tile  [D][8 chains] (2-way conflict on B fragments; 12 would be free) */
constexpr int kLPR = 12;       /* row pitch of the per-warp residual tile   [16 rows][8 chains]                     */

template <int D>
__host__ __device__ constexpr int l_stage_doubles() { return 8 * (D + 4) + 8; }          /* 8 padded rows + 8 labels */
template <int D>
__host__ __device__ constexpr int l_warp_doubles() { return D * kLPB + 16 * kLPR; }
template <int D>
__host__ __device__ constexpr size_t l_smem_bytes(int warps, int nst) {
    return sizeof(double) * ((size_t)nst * l_stage_doubles<D>() + (size_t)warps * l_warp_doubles<D>()) + 2 * (size_t)nst * sizeof(unsigned long long);
}

/* t = y eta - log1pexp(eta) and r = y - sigmoid(eta), sharing exp(-|eta|); same operations as the contract header */
__device__ __forceinline__ void logistic_terms(double eta, double y, double& t, double& r) {
    const double ex = amh::exp_(-fabs(eta));             /* in [0,1] */
    const double w = 1.0 + ex;
    /* both quotients have the denominator w in [1,2]: one shared reciprocal refinement + Markstein corrections
     * (amh_fastmath.cuh) = the correctly rounded IEEE quotients, without a slow-path branch.  The first numerator is
     * the rounding error of w: +0 or at least 2^-105 in magnitude (exact when w == 1), so nothing is subnormal. */
    double qc, sg0;
    div2_same_den(ex - (w - 1.0), 1.0, w, qc, sg0);
    const double l = (-amh::neglog_normal(w)) + qc;                          /* log_(w) for a normal w in [1,2] */
    const double l1p = (eta > 0.0 ? eta : 0.0) + l;
    const double sg = eta >= 0.0 ? sg0 : ex * sg0;
    t = y * eta - l1p;
    r = y - sg;
}

/* RW = true: the random-walk MH step (mh-core.jl:92-117) on the same target with a zero-mean proposal -- isotropic,
 * diagonal, or full covariance (candidate = x + L z as a small DMMA mat-vec, see phase A) -- the same kernel without the
 * gradient: candidate = x + sigma_j z_j (the per-thread kernel's two roundings),
 * GEMM1 + the log-likelihood terms only, Hastings term exactly 0.  RWMH on a logistic regression ran the generic
 * per-thread kernel before: 4.6e5 chain-steps/s on config 4's model, 9 x SLOWER than MALA on it (profiles/r2_dim_cliffs.txt). */
template <int D, bool RW = false>
__global__ void __launch_bounds__(512)
mala_logistic_kernel(const __grid_constant__ MalaLArgs a) {
    static_assert(D % 32 == 0 && D <= 128, "D/4 dims per lane part, D/8 noise blocks per part");
    constexpr int KT1 = D / 4;         /* k-tiles of GEMM1            */
    constexpr int MT2 = D / 8;         /* m-tiles (features) of GEMM2 */
    constexpr int XP = D + 4;          /* padded row pitch of a staged X block */
    constexpr int NPP = D / 8;         /* Philox blocks per lane part (4 parts per chain) */
    extern __shared__ __align__(16) double lsm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nwarps = (blockDim.x >> 5) - 1;                       /* consumer warps; the last warp is the TMA producer */
    const int kLStages = a.nst;
    double* stages = lsm;
    double* Bs = lsm + (size_t)kLStages * l_stage_doubles<D>() + (size_t)(warp < nwarps ? warp : 0) * l_warp_doubles<D>();   /* candidate tile [D][8] */
    double* Rs = Bs + D * kLPB;                                                                     /* residual tile [16][8] */
    unsigned long long* full = reinterpret_cast<unsigned long long*>(lsm + (size_t)kLStages * l_stage_doubles<D>() + (size_t)nwarps * l_warp_doubles<D>());
    unsigned long long* empty = full + kLStages;
    const long long pitch = a.st.pitch;
    const int fr = lane >> 2, fc = lane & 3;
    const int cl = lane & 7, part = lane >> 3;                      /* chain lane / quarter of the dimensions */
    const long long cbase = ((long long)blockIdx.x * nwarps + warp) * 8;
    const bool warp_active = cbase < a.st.n;                        /* idle warps still take part in the ring */
    const long long ch = cbase + cl;
    const bool active = ch < a.st.n;
    constexpr unsigned stage_bytes = (unsigned)(8 * D * sizeof(double) + 8 * sizeof(double));

    if (threadIdx.x == 0) {
        for (int s = 0; s < kLStages; ++s) { l_mbar_init(full + s, 1); l_mbar_init(empty + s, nwarps); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const long long total_blocks = (long long)a.nsteps * a.nblk;
    if (warp == nwarps) {
        /* ---- producer warp: keeps the X ring full; one elected lane drives the TMA engine ---- */
        if (lane == 0) {
            for (long long seq = 0; seq < total_blocks; ++seq) {
                const int blk = (int)(seq % a.nblk), st = (int)(seq % kLStages);
                if (seq >= kLStages) l_mbar_wait(empty + st, (unsigned)(((seq / kLStages) - 1) & 1));
                double* dst = stages + (size_t)st * l_stage_doubles<D>();
                l_mbar_expect_tx(full + st, stage_bytes);
#pragma unroll
                for (int r = 0; r < 8; ++r) l_bulk_g2s(dst + r * XP, a.Xp + ((size_t)blk * 8 + r) * D, (unsigned)(D * sizeof(double)), full + st);
                l_bulk_g2s(dst + 8 * XP, a.yp + (size_t)blk * 8, (unsigned)(8 * sizeof(double)), full + st);
            }
        }
        return;
    }
    const unsigned long long seed = active ? a.st.seeds[ch] : 0ull;
    double lp = active ? a.st.lp[ch] : 0.0;
    unsigned nacc = 0u;
    unsigned char accepted = active ? a.st.acc[ch] : (unsigned char)0;
    long long consumed = 0;                                         /* X blocks consumed over the whole launch */

    for (int s = 0; s < a.nsteps; ++s) {
        const unsigned long long k = a.step0 + (unsigned long long)s + 1ull;
        /* ---- phase A: candidate = x + (sigma z + drift grad)   (MALA.jl:70 -> proposal.jl:49-56) ---- */
        double e = 0.0;
        {
            double z[2 * NPP];
            if (a.st.cv == AMH_CONTRACT_V2) {
                /* contract v2: four normals per block -> D/16 blocks per lane part, the exponential in block D/4 */
                constexpr int NB2 = D / 16;
                const unsigned long long B2 = a.blocks_per_step;
                const unsigned long long b0 = k * B2 + (unsigned long long)(NB2 * part);
                if constexpr (NB2 > 4) {
                    noise_group<4, false, 2>(seed, b0, 0ull, z, e);
                    noise_group<NB2 - 4, true, 2>(seed, b0 + 4, k * B2 + a.exp_block, z + 16, e);
                } else {
                    noise_group<NB2, true, 2>(seed, b0, k * B2 + a.exp_block, z, e);
                }
            } else {
            const unsigned long long Bv1 = a.blocks_per_step;
            const unsigned long long b0 = k * Bv1 + (unsigned long long)(NPP * part);
            if constexpr (NPP > 8) {
                noise_group<8, false>(seed, b0, 0ull, z, e);
                noise_group<NPP - 8, true>(seed, b0 + 8, k * Bv1 + a.exp_block, z + 16, e);
            } else {
                noise_group<NPP, true>(seed, b0, k * Bv1 + a.exp_block, z, e);
            }
            }
            bool drawn = false;
            if constexpr (RW) {
                if (a.Lf) {
                    /* full-covariance proposal (proposal.jl:41-56 with MvNormal(0, Sigma)): candidate = x + L z as a DMMA
                     * mat-vec for the warp's 8 chains.  z goes into the candidate tile, the row blocks are produced from the
                     * last to the first and overwrite it in place (block mb reads rows <= 8 mb + 7 only); the accumulation over
                     * a row's tiles is the per-thread kernel's fma chain (the first product onto 0, zeros above the diagonal) */
                    drawn = true;
#pragma unroll
                    for (int i = 0; i < 2 * NPP; ++i) Bs[(2 * NPP * part + i) * kLPB + cl] = z[i];
                    __syncwarp();
                    const long long c0 = cbase + 2 * fc;
                    for (int mb = D / 8 - 1; mb >= 0; --mb) {
                        double y0 = 0.0, y1 = 0.0;
                        for (int kb = 0; kb <= 2 * mb + 1; ++kb) {
                            const double af = __ldg(a.Lf + ((size_t)(mb * (mb + 1) + kb)) * 32 + lane);
                            const double bf = Bs[(4 * kb + fc) * kLPB + fr];
                            l_dmma(y0, y1, af, bf);
                        }
                        __syncwarp();
                        const long long o = (long long)(8 * mb + fr) * pitch + c0;
                        const double x0 = (c0 < a.st.n) ? a.st.X[o] : 0.0;
                        const double x1 = (c0 + 1 < a.st.n) ? a.st.X[o + 1] : 0.0;
                        *reinterpret_cast<double2*>(Bs + (8 * mb + fr) * kLPB + 2 * fc) = make_double2(x0 + y0, x1 + y1);
                        __syncwarp();
                    }
                }
            }
            if (!drawn) {
#pragma unroll
            for (int i = 0; i < 2 * NPP; ++i) {
                const int j = 2 * NPP * part + i;
                const long long o = (long long)j * pitch + ch;
                double c = 0.0;
                if (active) {
                    /* the scale is 0 in the padding rows (and x, grad stay 0 there): c = 0 */
                    if constexpr (RW) {
                        c = a.st.X[o] + __ldg(a.dscale + j) * z[i];
                    } else {
                        c = a.st.X[o] + (__ldg(a.dscale + j) * z[i] + a.drift * a.st.G[o]);
                        a.Xc[o] = c;
                    }
                }
                Bs[j * kLPB + cl] = c;
            }
            }
        }
        __syncwarp();
        /* ---- phase B: stream X; eta = X c, residuals, gg = X' r ---- */
        double gg[MT2][2];
#pragma unroll
        for (int m = 0; m < MT2; ++m) { gg[m][0] = 0.0; gg[m][1] = 0.0; }
        double ll0 = 0.0, ll1 = 0.0;
        for (int blk = 0; blk < a.nblk; blk += 2, consumed += 2) {
            /* two 8-row blocks per iteration: their GEMM1 chains (32 dependent DMMAs each) interleave */
            const int st0 = (int)(consumed % kLStages), st1 = (int)((consumed + 1) % kLStages);
            l_mbar_wait(full + st0, (unsigned)((consumed / kLStages) & 1));
            l_mbar_wait(full + st1, (unsigned)(((consumed + 1) / kLStages) & 1));
            const double* Xs0 = stages + (size_t)st0 * l_stage_doubles<D>();
            const double* Xs1 = stages + (size_t)st1 * l_stage_doubles<D>();
            if (warp_active) {
                double e00 = 0.0, e01 = 0.0, e10 = 0.0, e11 = 0.0;
#pragma unroll
                for (int kb = 0; kb < KT1; ++kb) {
                    const double bf = Bs[(4 * kb + fc) * kLPB + fr];
                    const double af0 = Xs0[fr * XP + 4 * kb + fc];
                    const double af1 = Xs1[fr * XP + 4 * kb + fc];
                    l_dmma(e00, e01, af0, bf);
                    l_dmma(e10, e11, af1, bf);
                }
                const int row0 = blk * 8 + fr;
                const double y0 = Xs0[8 * XP + fr], y1 = Xs1[8 * XP + fr];
                double t00, r00, t01, r01, t10, r10, t11, r11;
                logistic_terms(e00, y0, t00, r00);
                logistic_terms(e01, y0, t01, r01);
                logistic_terms(e10, y1, t10, r10);
                logistic_terms(e11, y1, t11, r11);
                if (row0 < a.nrows) { ll0 = ll0 + t00; ll1 = ll1 + t01; }
                if (row0 + 8 < a.nrows) { ll0 = ll0 + t10; ll1 = ll1 + t11; }
                if constexpr (!RW) {
                *reinterpret_cast<double2*>(Rs + fr * kLPR + 2 * fc) = make_double2(r00, r01);
                *reinterpret_cast<double2*>(Rs + (8 + fr) * kLPR + 2 * fc) = make_double2(r10, r11);
                __syncwarp();
                const double rb0 = Rs[fc * kLPR + fr];                 /* B fragments: r[k = row][n = chain fr], 4 k-tiles */
                const double rb1 = Rs[(4 + fc) * kLPR + fr];
                const double rb2 = Rs[(8 + fc) * kLPR + fr];
                const double rb3 = Rs[(12 + fc) * kLPR + fr];
#pragma unroll
                for (int m = 0; m < MT2; ++m) {
                    const double a0 = Xs0[fc * XP + 8 * m + fr];       /* A fragments: X'[m = feature 8m+fr][k = row] */
                    const double a1 = Xs0[(4 + fc) * XP + 8 * m + fr];
                    const double a2 = Xs1[fc * XP + 8 * m + fr];
                    const double a3 = Xs1[(4 + fc) * XP + 8 * m + fr];
                    l_dmma(gg[m][0], gg[m][1], a0, rb0);
                    l_dmma(gg[m][0], gg[m][1], a1, rb1);
                    l_dmma(gg[m][0], gg[m][1], a2, rb2);
                    l_dmma(gg[m][0], gg[m][1], a3, rb3);
                }
                }
            }
            __syncwarp();
            if (lane == 0) { l_mbar_arrive(empty + st0); l_mbar_arrive(empty + st1); }
        }
        /* log-likelihood: tree over the 8 fragment rows (lane bits 2..4) */
#pragma unroll
        for (int m = 4; m <= 16; m <<= 1) {
            ll0 = ll0 + __shfl_xor_sync(0xffffffffu, ll0, m);
            ll1 = ll1 + __shfl_xor_sync(0xffffffffu, ll1, m);
        }
        /* gradient at the candidate in fragment layout: features 8m+fr, chains cbase + 2fc + {0,1} */
        if (!RW && warp_active) {
#pragma unroll
            for (int m = 0; m < MT2; ++m) {
                const int j = 8 * m + fr;
                const double2 cc = *reinterpret_cast<const double2*>(Bs + j * kLPB + 2 * fc);
                const double g0 = gg[m][0] - cc.x * a.invtau2;
                const double g1 = gg[m][1] - cc.y * a.invtau2;
                *reinterpret_cast<double2*>(a.Gc + (long long)j * pitch + cbase + 2 * fc) = make_double2(g0, g1);
            }
        }
        __syncwarp();
        /* ---- phase C: per chain (4 lanes carry the same chain): lp, Hastings terms, accept ---- */
        const double llsel0 = __shfl_sync(0xffffffffu, ll0, cl >> 1);
        const double llsel1 = __shfl_sync(0xffffffffu, ll1, cl >> 1);
        const double ll = (cl & 1) ? llsel1 : llsel0;
        double q = 0.0, A = 0.0, Bq = 0.0;
        if (active) {
            if constexpr (RW) {
                for (int j = 0; j < D; ++j) {
                    const double c = Bs[j * kLPB + cl];
                    q = (j == 0) ? c * c : fma(c, c, q);
                }
            } else {
            for (int j = 0; j < D; ++j) {
                const long long o = (long long)j * pitch + ch;
                const double c = Bs[j * kLPB + cl];
                const double xi = a.st.X[o], gi = a.st.G[o], gci = __ldcg(a.Gc + o);
                q = (j == 0) ? c * c : fma(c, c, q);
                const double da = (xi - c) - a.drift * gci;
                const double db = (c - xi) - a.drift * gi;
                A = (j == 0) ? da * da : fma(da, da, A);
                Bq = (j == 0) ? db * db : fma(db, db, Bq);
            }
            }
        }
        const double lp_c = ll - q * a.inv2tau2;
        const double logratio = RW ? 0.0 : (-0.5 * (A / a.sigma2)) - (-0.5 * (Bq / a.sigma2));
        const double loga = (lp_c - lp) + logratio;
        if (active && -e < loga) {                                   /* MALA.jl:86 */
#pragma unroll 4
            for (int i = 0; i < 2 * NPP; ++i) {
                const int j = 2 * NPP * part + i;
                const long long o = (long long)j * pitch + ch;
                a.st.X[o] = Bs[j * kLPB + cl];
                if constexpr (!RW) a.st.G[o] = __ldcg(a.Gc + o);
            }
            lp = lp_c;
            accepted = 1;
            ++nacc;
        } else {
            accepted = 0;
        }
        __syncwarp();
    }

    if (!active) return;
    if (a.sv.out || a.sv.sum) {
        for (int i = 0; i < 2 * NPP; ++i) {
            const int j = 2 * NPP * part + i;
            if (j >= a.d_real) break;                                  /* padding rows are not part of the sample */
            const long long o = (long long)j * pitch + ch;
            const double v = a.st.X[o];
            if (a.sv.out) a.sv.out[(long long)j * a.sv.out_pitch + ch] = v;
            if (a.sv.sum) {
                save_moments(a.sv, o, v);
            }
        }
    }
    if (part == 0) {
        a.st.lp[ch] = lp;
        a.st.nacc[ch] = a.st.nacc[ch] + (unsigned long long)nacc;
        a.st.acc[ch] = accepted;
        if (a.sv.out) a.sv.out[(long long)a.d_real * a.sv.out_pitch + ch] = lp;
        if (a.sv.acc_out) a.sv.acc_out[ch] = accepted;
    }
}

/* any dimension up to 128: the features are padded to 32 / 64 / 128 with zero columns of X (an exact no-op in every dot
 * product), zero noise scale and zero state rows, and the noise blocks are indexed with the real dimension's count */
static int logistic_padded_dim(int d) { return d <= 32 ? 32 : d <= 64 ? 64 : 128; }
bool mala_logistic_eligible(const amh_run& r) {
    const int d = r.dim;
    return r.sampler->d.kind == AMH_SAMPLER_MALA && r.target->kind == AMH_TARGET_LOGISTIC && d >= 1 && d <= 128 &&
           r.x_rows >= logistic_padded_dim(d) && r.target->ndata >= 64 && r.pitch % 32 == 0 &&
           (d % 32 == 0 || std::getenv("AMH_LOGISTIC_NO_PAD") == nullptr);
}

template <int D, bool RW = false>
static int launch_mala_logistic_t(amh_run& r, int nsteps, const SaveArgs& sv) {
    const amh_sampler& s = *r.sampler;
    const amh_target& t = *r.target;
    const long long n = t.ndata;
    const int d = r.dim;                                /* <= D: features d..D-1 are padding */
    const int nblk = (int)((n + 15) / 16) * 2;          /* 8-row blocks, consumed in pairs */
    const size_t np = (size_t)r.pitch;
    if (!r.scratch) {
        /* [Xpad | ypad | noise scales | MALA only: Xc | Gc] */
        const size_t nx = (size_t)nblk * 8 * D, ny = (size_t)nblk * 8;
        const size_t tail = (size_t)D + (RW ? 0 : 2 * (size_t)D * np);
        const int rca = dmalloc(r.ctx, &r.scratch, sizeof(double) * (nx + ny + tail));
        if (rca) return rca;
        double* base = (double*)r.scratch;
        AMH_CUDA_TRY(cudaMemsetAsync(base, 0, sizeof(double) * (nx + ny + tail), r.ctx->stream));
        AMH_CUDA_TRY(cudaMemcpy2DAsync(base, sizeof(double) * D, t.dblob + 1, sizeof(double) * d, sizeof(double) * d, (size_t)n,
                                       cudaMemcpyDeviceToDevice, r.ctx->stream));
        AMH_CUDA_TRY(cudaMemcpyAsync(base + nx, t.dblob + 1 + (size_t)n * d, sizeof(double) * (size_t)n, cudaMemcpyDeviceToDevice, r.ctx->stream));
        std::vector<double> sc(D, 0.0);
        const bool fullcov = RW && s.d.cov_kind == AMH_COV_FULL;
        for (int i = 0; i < d && !fullcov; ++i) sc[i] = !RW ? s.mala_sigma : s.d.cov_kind == AMH_COV_DIAG ? s.scale[i] : s.scale[0];
        AMH_CUDA_TRY(cudaMemcpyAsync(base + nx + ny, sc.data(), sizeof(double) * D, cudaMemcpyHostToDevice, r.ctx->stream));
        std::vector<double> lf;
        if (fullcov) {
            /* A fragments of the proposal's factor, zero-padded to D x D (the layout of amh_launch_mh_tc.cu's build_frags) */
            constexpr int NBL = D / 8;
            lf.assign((size_t)NBL * (NBL + 1) * 32, 0.0);
            for (int mb = 0; mb < NBL; ++mb)
                for (int kb = 0; kb <= 2 * mb + 1; ++kb)
                    for (int ln = 0; ln < 32; ++ln) {
                        const int row = 8 * mb + ln / 4, col = 4 * kb + ln % 4;
                        if (col <= row && row < d) lf[((size_t)(mb * (mb + 1) + kb)) * 32 + ln] = s.scale[tri_h(row, col)];
                    }
            const int rcb = dmalloc(r.ctx, &r.scratch2, sizeof(double) * lf.size());
            if (rcb) return rcb;
            AMH_CUDA_TRY(cudaMemcpyAsync(r.scratch2, lf.data(), sizeof(double) * lf.size(), cudaMemcpyHostToDevice, r.ctx->stream));
        }
        AMH_CUDA_TRY(sync_stream(r.ctx, r.ctx->stream));        /* `sc`, `lf` are stack temporaries */
    }
    MalaLArgs a;
    std::memset(&a, 0, sizeof(a));
    a.st = chain_state(r);
    a.sv = sv;
    a.nsteps = nsteps;
    a.step0 = (unsigned long long)r.step;
    a.sigma = s.mala_sigma; a.sigma2 = s.d.mala_sigma2; a.drift = s.d.mala_drift;
    double* base = (double*)r.scratch;
    a.Xp = base;
    a.yp = base + (size_t)nblk * 8 * D;
    a.dscale = base + (size_t)nblk * 8 * D + (size_t)nblk * 8;
    a.Xc = RW ? nullptr : const_cast<double*>(a.dscale) + D;
    a.Gc = RW ? nullptr : a.Xc + (size_t)D * np;
    a.d_real = d;
    a.Lf = (RW && s.d.cov_kind == AMH_COV_FULL) ? (const double*)r.scratch2 : nullptr;
    a.exp_block = (unsigned long long)(r.cv == AMH_CONTRACT_V2 ? (d + 3) / 4 : (d + 1) / 2);
    a.blocks_per_step = a.exp_block + 1ull;
    a.nrows = n;
    a.nblk = nblk;
    a.inv2tau2 = t.inv2tau2; a.invtau2 = t.invtau2;
    /* one CTA per SM: W consumer warps (8 chains each) + 1 producer warp; W chosen so that the grid fills the SMs evenly,
     * the rest of the shared memory becomes X-ring stages (prefetch depth hides the TMA round trip) */
    const long long groups = (r.n + 7) / 8;
    const int sms = r.ctx->sm_count;
    int warps = 14;
    {
        /* the largest W among the most even fillings; runs with few chains end up with small CTAs on many SMs (1 024 chains
         * of config 4's model: 9 CTAs of 15 warps before, 128 CTAs of one consumer warp now -- every CTA streams X for
         * itself, from L2) */
        double best_eff = 0;
        for (int w = 15; w >= 1; --w) {
            if (l_smem_bytes<D>(w, 4) + 1024 > 227 * 1024) continue;
            const long long ctas = (groups + w - 1) / w;
            const long long waves = (ctas + sms - 1) / sms;
            const double eff = (double)groups / ((double)waves * sms * w);
            if (eff > best_eff + 1e-9) { best_eff = eff; warps = w; }
        }
        if (const char* ev = std::getenv("AMH_K3L_WARPS")) {                          /* A/B and test switch */
            const int w = std::atoi(ev);
            if (w >= 1 && w <= 15 && l_smem_bytes<D>(w, 4) + 1024 <= 227 * 1024) warps = w;
        }
    }
    int nst = 4;
    while (nst + 2 <= 16 && l_smem_bytes<D>(warps, nst + 2) + 1024 <= 227 * 1024) nst += 2;
    a.nst = nst;
    const size_t smem = l_smem_bytes<D>(warps, nst);
    auto kern = mala_logistic_kernel<D, RW>;
    AMH_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const unsigned grid = (unsigned)((groups + warps - 1) / warps);
    kern<<<grid, 32 * (warps + 1), smem, r.ctx->stream>>>(a);
    AMH_CUDA_TRY(cudaGetLastError());
    r.launches += 1;
    r.pending_launches += 1;
    return AMH_OK;
}

int launch_mala_logistic(amh_run& r, int nsteps, const SaveArgs& sv) {
    switch (logistic_padded_dim(r.dim)) {
    case 32: return launch_mala_logistic_t<32>(r, nsteps, sv);
    case 64: return launch_mala_logistic_t<64>(r, nsteps, sv);
    case 128: return launch_mala_logistic_t<128>(r, nsteps, sv);
    }
    return fail(AMH_ERR_INVALID, "tiled logistic MALA: unsupported dimension");
}

/* RWMH on the many-row logistic target with an isotropic / diagonal zero-mean proposal: the RW variant of the kernel */
bool mh_logistic_eligible(const amh_run& r) {
    const amh_sampler& s = *r.sampler;
    const int d = r.dim;
    if (s.d.kind != AMH_SAMPLER_RW || r.target->kind != AMH_TARGET_LOGISTIC) return false;
    if (d < 1 || d > 128 || r.x_rows < logistic_padded_dim(d) || r.target->ndata < 64 || r.pitch % 32) return false;
    if (d % 32 != 0 && std::getenv("AMH_LOGISTIC_NO_PAD")) return false;
    if (s.has_mean || s.by_components()) return false;
    if (s.d.cov_kind != AMH_COV_DIAG && s.d.cov_kind != AMH_COV_SCALAR && s.d.cov_kind != AMH_COV_FULL) return false;
    return std::getenv("AMH_MH_NO_LOGISTIC") == nullptr;             /* A/B switch: the generic per-thread kernel */
}
int launch_mh_logistic(amh_run& r, int nsteps, const SaveArgs& sv) {
    switch (logistic_padded_dim(r.dim)) {
    case 32: return launch_mala_logistic_t<32, true>(r, nsteps, sv);
    case 64: return launch_mala_logistic_t<64, true>(r, nsteps, sv);
    case 128: return launch_mala_logistic_t<128, true>(r, nsteps, sv);
    }
    return fail(AMH_ERR_INVALID, "tiled logistic RWMH: unsupported dimension");
}

}  // namespace amhh
