/* amh_launch_mh_tc.cu -- K1T: the random-walk / static MH step with both triangular mat-vecs
 * (proposal L z, target U (c - mu)) on the FP64 tensor cores (mma.sync m8n8k4 = DMMA).
 *
 * Why (tools/ubench/dmma_probe.cu, measured on B200): DMMA and DFMA share ONE FP64 datapath of
 * 64 FMA/clk/SM, so tensor cores add no FP64 throughput -- but one DMMA retires 256 FMAs for one
 * issue slot and takes its L/U operand from a register fragment, where the per-thread DFMA form
 * needs 1056 DFMA + 528 constant loads per chain-step, 64+ live registers for z and c, and ~70 KB
 * of straight-line code.  The warp is the natural tile: its 32 chains are the N dimension,
 *     Y[32 x 32 chains] = L[32 x 32] * Z[32 x 32 chains].
 * DMMA accumulates k = 0..3 as a sequential IEEE fma chain (verified by the probe), so the result is
 * bit-identical to the contract's row dot products (include/amh_contract.h, oracle `draw` / `logp`).
 *
 * Per warp and MCMC step (mh-core.jl:92-117, proposal.jl:41-56):
 *   phase 0  every lane = one chain: batched Philox / Box-Muller -> Z[k][lane] in shared memory
 *   phase 1  row blocks 3..0: DMMA chains over the lower-triangular tiles of L; c = x + y in the
 *            accumulator layout (x is re-read from global/L2: 528 B per chain-step is exactly the
 *            algorithmic traffic); C overwrites Z in place (descending row blocks make that safe)
 *   phase 2  row blocks 0..3: W = U (C - mu) by DMMA; each 8 x 32 slab of W goes through a 2.5 KB
 *            staging tile back to the chain lanes, which accumulate q = sum w_i^2 in index order
 *   accept   chain lanes: -randexp < lp_c - lp (strict); accepted lanes copy their column of C to X.
 * Shared memory: 11.6 KB per warp -> 14 warps/SM = the whole 65 536-chain problem in one wave.
 */
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include "amh_params.cuh"

namespace amhh {
using namespace amhd;

struct MhTcArgs {
    ChainState st;
    SaveArgs sv;
    int nsteps;
    int is_rw;
    int mu_zero;
    unsigned long long step0;
    const double* Lf;          /* [NT][32] A fragments of the proposal factor   */
    const double* Uf;          /* [NT][32] A fragments of the target factor     */
    const double* mu;          /* [D]                                           */
    double c0;
    const double* dscale;      /* [D] standard deviations of a diagonal / isotropic proposal (COVD variants) */
    int pace;                  /* nanoseconds of optional pause per step (see the step loop of K1T16) */
    /* PAD variants (amh_launch_mh_tcp.cu): the run's dimension d <= D, its noise blocks per step and the offset of the
     * exponential's block inside a step (contract v2: ceil(d/4) + 1 and ceil(d/4)) */
    int d_real;
    unsigned long long blocks_per_step, exp_block;
};

__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
        : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

/* x fragment load pinned in program order (volatile): ptxas otherwise hoists all of them to the top of the
 * step and then spills the values, which exposes the full L2 latency on the spill store (ncu, round 1) */
__device__ __forceinline__ double2 ld_x_frag(const double* p) {
    double2 v;
    asm volatile("ld.global.cg.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
    return v;
}

/* A-fragment load pinned in program order (explicit software pipeline, see ld_x_frag) */
__device__ __forceinline__ double ld_a_frag(const double* p) {
    double v;
    asm volatile("ld.global.nc.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
}
/* flattened lower-triangular tile sequences: row block mb has k-tiles 0 .. 2mb+1 */
__host__ __device__ constexpr int tseq_mb(int nb, int t, bool desc) {
    for (int i = 0; i < nb; ++i) {
        const int mb = desc ? nb - 1 - i : i;
        const int cnt = 2 * mb + 2;
        if (t < cnt) return mb;
        t -= cnt;
    }
    return -1;
}
__host__ __device__ constexpr int tseq_kb(int nb, int t, bool desc) {
    for (int i = 0; i < nb; ++i) {
        const int mb = desc ? nb - 1 - i : i;
        const int cnt = 2 * mb + 2;
        if (t < cnt) return t;
        t -= cnt;
    }
    return -1;
}

/* Interleaved schedule: row blocks are processed in PAIRS (descending (NB-1,NB-2),(NB-3,NB-4).. or ascending
 * (0,1),(2,3)..), the k-tiles of the two blocks alternating, so that a warp always has 4 independent DMMA
 * accumulation chains (2 row blocks x 2 n-tiles) in flight instead of 2.  psched(NB, t, desc, what):
 * what = 0 -> row block of issue slot t, 1 -> k-tile, 2 -> 1 if t is the last slot of its pair. */
__host__ __device__ constexpr int psched(int nb, int t, bool desc, int what) {
    for (int p = 0; 2 * p < nb; ++p) {
        const int A = desc ? nb - 1 - 2 * p : 2 * p;
        const int Bq = desc ? nb - 2 - 2 * p : 2 * p + 1;
        const bool hasB = Bq >= 0 && Bq < nb;
        const int cA = 2 * A + 2, cB = hasB ? 2 * Bq + 2 : 0;
        if (t < cA + cB) {
            int c = 0;
            const int mx = cA > cB ? cA : cB;
            for (int k = 0; k < mx; ++k) {
                if (k < cA) { if (c == t) return what == 0 ? A : what == 1 ? k : (t == cA + cB - 1); ++c; }
                if (k < cB) { if (c == t) return what == 0 ? Bq : what == 1 ? k : (t == cA + cB - 1); ++c; }
            }
        }
        t -= cA + cB;
    }
    return -1;
}

/* the same schedule as compile-time tables: for more than four row blocks (padded dimensions above 32) folding the psched()
 * calls of the unrolled tile loops costs the compiler tens of minutes */
template <int NB, bool DESC>
struct PSchedTab {
    static constexpr int NT = NB * (NB + 1);
    int mb[NT], kb[NT], last[NT];
    constexpr PSchedTab() : mb{}, kb{}, last{} {
        for (int t = 0; t < NT; ++t) {
            mb[t] = psched(NB, t, DESC, 0);
            kb[t] = psched(NB, t, DESC, 1);
            last[t] = psched(NB, t, DESC, 2);
        }
    }
};

constexpr int kPZ = 36;        /* row pitch (doubles) of the Z/C tile: conflict-free B-fragment loads */
constexpr int kPW = 40;        /* row pitch of the W staging tile: conflict-free 128-bit stores       */

template <int D>
__host__ __device__ constexpr int tc_smem_doubles_per_warp() { return D * kPZ + 8 * kPW; }

template <int D, int WARPS>
__global__ void __launch_bounds__(32 * WARPS, 14 / WARPS)
mh_step_tc_kernel(const __grid_constant__ MhTcArgs a) {
    static_assert(D % 8 == 0 && D >= 8 && D <= 32, "row blocks of 8");
    constexpr int NB = D / 8;
    extern __shared__ double smem[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    double* __restrict__ ZC = smem + warp * tc_smem_doubles_per_warp<D>();
    double* __restrict__ WB = ZC + D * kPZ;
    const long long cbase = ((long long)blockIdx.x * WARPS + warp) * 32;     /* first chain of this warp */
    if (cbase >= a.st.n) return;                                              /* whole warp idle */
    const long long ch = cbase + lane;
    const bool active = ch < a.st.n;
    const long long pitch = a.st.pitch;
    const int fr = lane >> 2, fc = lane & 3;                                  /* fragment row / column ids */
    double* __restrict__ X = a.st.X;

    const unsigned long long seed = active ? a.st.seeds[ch] : 0ull;
    double lp = active ? a.st.lp[ch] : 0.0;
    unsigned long long nacc = active ? a.st.nacc[ch] : 0ull;
    unsigned char accepted = active ? a.st.acc[ch] : (unsigned char)0;
    constexpr unsigned long long B = (unsigned long long)((D + 1) / 2 + 1);

    for (int s = 0; s < a.nsteps; ++s) {
        const unsigned long long k = a.step0 + (unsigned long long)s + 1ull;
        /* ---- phase 0: noise of this step, one chain per lane -> Z[k][lane] ---- */
        double e;
        {
            double z[D];
            step_noise_fixed<D>(seed, k * B, z, e);
#pragma unroll
            for (int i = 0; i < D; ++i) ZC[i * kPZ + lane] = z[i];
        }
        /* x in accumulator-fragment layout: rows 8mb+fr, chains cbase + 8nb + 2fc + {0,1} (L2 resident) */
        double2 xf[NB][4];
        if (a.is_rw) {
#pragma unroll
            for (int mb = 0; mb < NB; ++mb)
#pragma unroll
                for (int nb = 0; nb < 4; ++nb)
                    xf[mb][nb] = __ldcg(reinterpret_cast<const double2*>(X + (long long)(8 * mb + fr) * pitch + cbase + 8 * nb + 2 * fc));
        }
        __syncwarp();
        /* ---- phase 1: C = X + L Z, row blocks in descending order, in place ---- */
#pragma unroll
        for (int mbi = 0; mbi < NB; ++mbi) {
            const int mb = NB - 1 - mbi;
            double acc[4][2];
#pragma unroll
            for (int nb = 0; nb < 4; ++nb) { acc[nb][0] = 0.0; acc[nb][1] = 0.0; }
#pragma unroll
            for (int kb = 0; kb <= 2 * mb + 1; ++kb) {
                const double af = __ldg(a.Lf + (mb * (mb + 1) + kb) * 32 + lane);
#pragma unroll
                for (int nb = 0; nb < 4; ++nb) {
                    const double bf = ZC[(4 * kb + fc) * kPZ + 8 * nb + fr];
                    dmma(acc[nb][0], acc[nb][1], af, bf);
                }
            }
#pragma unroll
            for (int nb = 0; nb < 4; ++nb) {
                double2 c;
                if (a.is_rw) { c.x = xf[mb][nb].x + acc[nb][0]; c.y = xf[mb][nb].y + acc[nb][1]; }
                else { c.x = acc[nb][0]; c.y = acc[nb][1]; }
                *reinterpret_cast<double2*>(ZC + (8 * mb + fr) * kPZ + 8 * nb + 2 * fc) = c;
            }
        }
        __syncwarp();
        /* ---- phase 2: W = U (C - mu), q = sum_i w_i^2 in index order ---- */
        double q = 0.0;
#pragma unroll
        for (int mb = 0; mb < NB; ++mb) {
            double acc[4][2];
#pragma unroll
            for (int nb = 0; nb < 4; ++nb) { acc[nb][0] = 0.0; acc[nb][1] = 0.0; }
#pragma unroll
            for (int kb = 0; kb <= 2 * mb + 1; ++kb) {
                const double af = __ldg(a.Uf + (mb * (mb + 1) + kb) * 32 + lane);
                const double muk = a.mu_zero ? 0.0 : __ldg(a.mu + 4 * kb + fc);
#pragma unroll
                for (int nb = 0; nb < 4; ++nb) {
                    const double cv = ZC[(4 * kb + fc) * kPZ + 8 * nb + fr];
                    const double bf = a.mu_zero ? cv : cv - muk;
                    dmma(acc[nb][0], acc[nb][1], af, bf);
                }
            }
#pragma unroll
            for (int nb = 0; nb < 4; ++nb)
                *reinterpret_cast<double2*>(WB + fr * kPW + 8 * nb + 2 * fc) = make_double2(acc[nb][0], acc[nb][1]);
            __syncwarp();
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                const double w = WB[r * kPW + lane];
                q = (mb == 0 && r == 0) ? w * w : fma(w, w, q);
            }
            __syncwarp();
        }
        const double lp_c = fma(-0.5, q, a.c0);
        /* ---- accept / reject (mh-core.jl:104-114); the Hastings term is exactly 0 on this path ---- */
        const double loga = (lp_c - lp) + 0.0;
        if (active && -e < loga) {
#pragma unroll
            for (int i = 0; i < D; ++i) X[(long long)i * pitch + ch] = ZC[i * kPZ + lane];
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
#pragma unroll 4
        for (int i = 0; i < D; ++i) {
            const long long o = (long long)i * pitch + ch;
            const double v = X[o];
            if (a.sv.out) a.sv.out[(long long)i * a.sv.out_pitch + ch] = v;
            if (a.sv.sum) {
                save_moments(a.sv, o, v);
            }
        }
    }
    a.st.lp[ch] = lp;
    a.st.nacc[ch] = nacc;
    a.st.acc[ch] = accepted;
    if (a.sv.out) a.sv.out[(long long)D * a.sv.out_pitch + ch] = lp;
    if (a.sv.acc_out) a.sv.acc_out[ch] = accepted;
}

/* ---------------------------------------------------------------------------
 * K1T16: same algorithm, 16 chains per warp.  Two lanes share a chain: lane = (half, cl); each half
 * generates 8 of the 16 Philox/Box-Muller blocks of the chain's step, the DMMA tiles cover N = 16
 * chains (2 n-tiles), and both lanes of a chain carry the same q / lp / accept decision (no shuffles).
 * Twice the warps for the same problem: 28 resident warps per SM instead of 14, which is what hides
 * the integer / FP64 dependency latency of the noise phase (ncu: issue slots 34% busy with 14 warps). */
constexpr int kPZ16 = 20;
constexpr int kPW16 = 24;
template <int D>
__host__ __device__ constexpr int tc16_smem_doubles_per_warp() { return D * kPZ16 + 8 * kPW16; }

/* WARPS: 28 = ONE CTA per SM (default), 4 = seven CTAs per SM.  Same code per warp, same 28 resident warps -- but the
 * warp schedulers arbitrate by warp slot (B200: highest slot first, round-robin among equals): the warps of seven
 * small CTAs get equal treatment and stay in lock-step, all in the noise phase (issue-slot and DFMA bound) or all in the
 * DMMA phase (FP64-pipe bound, few issue slots) at the same time; the 28 warps of one CTA have distinct priorities,
 * drift apart and overlap the two phases.  Measured on C2: 4.93e9 -> 5.53e9 chain-steps/s (tools/bench_configs.py with
 * AMH_TC_WARPS=4 / 28).  A bound on the drift (no warp more than k steps ahead of the slowest) was tried to cut the
 * end-of-launch tail and LOSES throughput for every k (5.28e9 at k <= 16). */
/* COVD: the proposal covariance is diagonal (ScalMat / PDiagMat): v_i = sigma_i z_i needs no mat-vec, phase 1 is
 * c = x + sigma_i z_i on the chain lanes (the contract's two roundings, proposal.jl:41-56 with a diagonal factor). */
/* PAD (amh_launch_mh_tcp.cu, contract v2): the run's dimension is a.d_real <= D.  L, U, mu and the scales are padded with
 * zeros (rows / columns d_real..D-1), the state has D rows on the device (the padding rows stay 0), and the noise blocks
 * are indexed with the real dimension's blocks per step, so every real coordinate sees exactly the contract's arithmetic:
 * a padding column adds fma(0, z, acc) = acc to a row's dot product -- what the zeros above the diagonal inside the
 * diagonal tiles do already -- and a padding row contributes fma(0, 0, q) = q to the quadratic form.  The normals the
 * lanes draw for padding rows come from blocks of the chain's stream outside this step's range; they are finite and only
 * ever multiply zeros.  D up to 64 (fewer warps per SM: the Z / C tile grows with D). */
template <int D, int WARPS, bool MU_ZERO, bool IS_RW, bool COVD = false, int CV = 1, bool PAD = false>
__global__ void __launch_bounds__(32 * WARPS, (WARPS <= 7 ? 28 / WARPS : 1))
mh_step_tc16_kernel(const __grid_constant__ MhTcArgs a) {

    static_assert(D % 8 == 0 && D >= 8 && (D <= 32 || (PAD && D <= 128)), "row blocks of 8; D/4 (v1) or D/8 (v2) noise blocks per lane half");
    static_assert(!PAD || CV == 2, "padded dimensions: contract v2 only");
    constexpr int NB = D / 8;
    constexpr int NPB = (CV == 2) ? 4 : 2;     /* normals per Philox block (contract v1 / v2) */
    constexpr int NPH = D / (2 * NPB);         /* Philox blocks per half-chain lane */
    constexpr int HR = D / 2;                  /* rows of Z / X owned by a half      */
    /* tile schedule of the two mat-vecs: psched() folded by the optimiser up to D = 32 (the kernels bench.py times keep the
     * code they were tuned with), compile-time tables above */
    constexpr bool WIDE = D > 32;
    constexpr PSchedTab<(WIDE ? NB : 1), true> pst1{};
    constexpr PSchedTab<(WIDE ? NB : 1), false> pst2{};
#define PS1(t_, w_) (WIDE ? ((w_) == 0 ? pst1.mb[(WIDE ? (t_) : 0)] : (w_) == 1 ? pst1.kb[(WIDE ? (t_) : 0)] : pst1.last[(WIDE ? (t_) : 0)]) : psched(NB, (t_), true, (w_)))
#define PS2(t_, w_) (WIDE ? ((w_) == 0 ? pst2.mb[(WIDE ? (t_) : 0)] : (w_) == 1 ? pst2.kb[(WIDE ? (t_) : 0)] : pst2.last[(WIDE ? (t_) : 0)]) : psched(NB, (t_), false, (w_)))
    extern __shared__ double smem[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    double* __restrict__ ZC = smem + warp * tc16_smem_doubles_per_warp<D>();
    double* __restrict__ WB = ZC + D * kPZ16;
    const long long cbase = ((long long)blockIdx.x * WARPS + warp) * 16;
    if (cbase >= a.st.n) return;
    const int cl = lane & 15, half = lane >> 4;
    const long long ch = cbase + cl;
    const bool active = ch < a.st.n;
    const long long pitch = a.st.pitch;
    const int fr = lane >> 2, fc = lane & 3;
    double* __restrict__ X = a.st.X;

    const unsigned long long seed = active ? a.st.seeds[ch] : 0ull;
    double lp = active ? a.st.lp[ch] : 0.0;
    unsigned nacc = 0u;                         /* accepted moves of this launch */
    unsigned char accepted = active ? a.st.acc[ch] : (unsigned char)0;
    const unsigned long long B = PAD ? a.blocks_per_step : (unsigned long long)(D / NPB + 1);
    const unsigned long long EB = PAD ? a.exp_block : (unsigned long long)(D / NPB);      /* the exponential's block within a step */
    double e_next = 0.0;

    for (int s = 0; s < a.nsteps; ++s) {
        const unsigned long long k = a.step0 + (unsigned long long)s + 1ull;
        /* Optional pacing (AMH_TC_PACE nanoseconds, default 0 = never taken).  The untaken branch is kept on purpose: with
         * it at the top of the step ptxas schedules the whole loop body differently, and that schedule is 8 % faster
         * (C2: 5.18e9 -> 5.60e9 chain-steps/s, same box, A/B of two builds, profiles/r1_k1t16_cta_shape_ab.txt); an
         * asm memory clobber or a __syncwarp at the same place do not have this effect. */
        /* (contract v1 only: the v2 kernel -- the default -- is 1.6 % FASTER without the branch, 6.39e9 vs 6.28e9 chain-steps/s,
         * profiles/r2_k1t16_v2_ab.txt, so the default path no longer depends on this scheduling accident) */
        if constexpr (CV == 1) {
            if (a.pace > 0) __nanosleep((unsigned)a.pace);
        }
        double e;
        if constexpr (PAD) {
            /* groups of at most two blocks (four Box-Muller pairs in lock-step), the last one carries the exponential */
            const unsigned long long b0 = k * B + (unsigned long long)(NPH * half);
            double* zt = ZC + (HR * half) * kPZ16 + cl;
            constexpr int NG = (NPH + 1) / 2;
            double eh = 0.0;
#pragma unroll
            for (int g = 0; g < NG; ++g) {
                const bool last = g == NG - 1;
                if (!last) {
                    double dmy;
                    noise_group<2, false, CV>(seed, b0 + 2 * g, 0ull, zt + NPB * 2 * g * kPZ16, dmy, amh::amh_log_tab_dev, kPZ16);
                } else if ((s & 1) == 0) {
                    noise_group<NPH - 2 * (NG - 1), true, CV>(seed, b0 + 2 * (NG - 1), (k + (unsigned long long)half) * B + EB,
                                                              zt + NPB * 2 * (NG - 1) * kPZ16, eh, amh::amh_log_tab_dev, kPZ16);
                } else {
                    double dmy;
                    noise_group<NPH - 2 * (NG - 1), false, CV>(seed, b0 + 2 * (NG - 1), 0ull, zt + NPB * 2 * (NG - 1) * kPZ16, dmy,
                                                               amh::amh_log_tab_dev, kPZ16);
                }
            }
            if ((s & 1) == 0) {
                e = __shfl_sync(0xffffffffu, eh, cl);
                e_next = __shfl_sync(0xffffffffu, eh, cl + 16);
            } else {
                e = e_next;
            }
        } else {
            const unsigned long long b0 = k * B + (unsigned long long)(NPH * half);
            double* zt = ZC + (HR * half) * kPZ16 + cl;            /* Z[HR half + j][cl] */
            /* two lock-step batches when there are >= 6 pairs (v2, d = 32: 2 + 2 blocks; one batch of 4 blocks measured 2.5 % slower) */
            constexpr int G1 = (NPH * NPB >= 12) ? NPH / 2 : 0;
            if constexpr (G1 > 0) noise_group<(G1 > 0 ? G1 : 1), false, CV>(seed, b0, 0ull, zt, e, amh::amh_log_tab_dev, kPZ16);
            /* the exponential: both lanes of a chain run the same instructions, so on even steps of the launch lane half
             * h draws the exponential of step k + h, and odd steps draw none */
            if ((s & 1) == 0) {
                double eh;
                noise_group<NPH - G1, true, CV>(seed, b0 + G1, (k + (unsigned long long)half) * B + EB,
                                                zt + NPB * G1 * kPZ16, eh, amh::amh_log_tab_dev, kPZ16);
                e = __shfl_sync(0xffffffffu, eh, cl);
                e_next = __shfl_sync(0xffffffffu, eh, cl + 16);
            } else {
                double dummy;
                noise_group<NPH - G1, false, CV>(seed, b0 + G1, 0ull, zt + NPB * G1 * kPZ16, dummy, amh::amh_log_tab_dev, kPZ16);
                e = e_next;
            }
        }
        if constexpr (COVD) {
            /* every lane owns rows HR*half .. of its chain: the same tile entries it has just written */
#pragma unroll
            for (int i = 0; i < HR; ++i) {
                const int row = HR * half + i;
                const double t = __ldg(a.dscale + row) * ZC[row * kPZ16 + cl];
                double c = t;
                if (IS_RW) c = __ldcg(X + (long long)row * pitch + ch) + t;
                ZC[row * kPZ16 + cl] = c;
            }
        } else {
        /* x in accumulator-fragment layout, software-pipelined two row blocks ahead of its use */
        double2 xf[NB][2];
        const double* xp = X + (long long)fr * pitch + cbase + 2 * fc;
        if (IS_RW) {
#pragma unroll
            for (int mb = NB - 1; mb >= (NB >= 2 ? NB - 2 : 0); --mb)
#pragma unroll
                for (int nb = 0; nb < 2; ++nb) xf[mb][nb] = ld_x_frag(xp + (long long)(8 * mb) * pitch + 8 * nb);
        }
        __syncwarp();
        {
            constexpr int NT = NB * (NB + 1);
            constexpr int P = 4;                       /* A-fragment prefetch distance (tiles) */
            double aq[NT];
            double acc[NB][2][2];
#pragma unroll
            for (int j = 0; j < P && j < NT; ++j) {
                const int mbj = PS1(j, 0), kbj = PS1(j, 1);
                aq[j] = ld_a_frag(a.Lf + (mbj * (mbj + 1) + kbj) * 32 + lane);
            }
#pragma unroll
            for (int t = 0; t < NT; ++t) {
                const int mb = PS1(t, 0), kb = PS1(t, 1);
                if (t + P < NT) {
                    const int mbj = PS1(t + P, 0), kbj = PS1(t + P, 1);
                    aq[t + P] = ld_a_frag(a.Lf + (mbj * (mbj + 1) + kbj) * 32 + lane);
                }
                if (kb == 0) { acc[mb][0][0] = 0.0; acc[mb][0][1] = 0.0; acc[mb][1][0] = 0.0; acc[mb][1][1] = 0.0; }
#pragma unroll
                for (int nb = 0; nb < 2; ++nb) {
                    const double bf = ZC[(4 * kb + fc) * kPZ16 + 8 * nb + fr];
                    dmma(acc[mb][nb][0], acc[mb][nb][1], aq[t], bf);
                }
                if (PS1(t, 2) == 1) {
                    /* both row blocks of the pair have read Z: their rows of C may now replace it.  The reads feed
                     * mma.sync.aligned instructions that every lane has executed before it gets here, so they have
                     * completed warp-wide; the explicit warp barrier states that ordering for the memory model and for
                     * racecheck (which reported a warning-level write-after-read hazard without it) and costs nothing
                     * measurable: 5.641 / 5.643e9 without, 5.635 / 5.641e9 chain-steps/s with (profiles/r2_sanitizers.md) */
                    __syncwarp();
#pragma unroll
                    for (int u = 0; u < 2; ++u) {
                        const int mw = (u == 0) ? mb : ((mb & 1) == (NB & 1) ? mb + 1 : mb - 1);   /* the pair's other block */
                        const int pa = (NB - 1 - mb) / 2;                                          /* pair index (descending) */
                        const int hi = NB - 1 - 2 * pa, lo = NB - 2 - 2 * pa;
                        const int mq = (u == 0) ? hi : lo;
                        (void)mw;
                        if (mq >= 0) {
#pragma unroll
                            for (int nb = 0; nb < 2; ++nb) {
                                double2 c;
                                if (IS_RW) { c.x = xf[mq][nb].x + acc[mq][nb][0]; c.y = xf[mq][nb].y + acc[mq][nb][1]; }
                                else { c.x = acc[mq][nb][0]; c.y = acc[mq][nb][1]; }
                                *reinterpret_cast<double2*>(ZC + (8 * mq + fr) * kPZ16 + 8 * nb + 2 * fc) = c;
                            }
                            if (IS_RW && mq >= 2) {
#pragma unroll
                                for (int nb = 0; nb < 2; ++nb) xf[mq - 2][nb] = ld_x_frag(xp + (long long)(8 * (mq - 2)) * pitch + 8 * nb);
                            }
                        }
                    }
                }
            }
        }
        }   /* !COVD */
        __syncwarp();
        double q = 0.0;
        {
            constexpr int NT = NB * (NB + 1);
            constexpr int P = 4;
            double aq[NT];
            double acc[NB][2][2];
#pragma unroll
            for (int j = 0; j < P && j < NT; ++j) {
                const int mbj = PS2(j, 0), kbj = PS2(j, 1);
                aq[j] = ld_a_frag(a.Uf + (mbj * (mbj + 1) + kbj) * 32 + lane);
            }
#pragma unroll
            for (int t = 0; t < NT; ++t) {
                const int mb = PS2(t, 0), kb = PS2(t, 1);
                if (t + P < NT) {
                    const int mbj = PS2(t + P, 0), kbj = PS2(t + P, 1);
                    aq[t + P] = ld_a_frag(a.Uf + (mbj * (mbj + 1) + kbj) * 32 + lane);
                }
                if (kb == 0) { acc[mb][0][0] = 0.0; acc[mb][0][1] = 0.0; acc[mb][1][0] = 0.0; acc[mb][1][1] = 0.0; }
                double muk = 0.0;
                if (!MU_ZERO) muk = __ldg(a.mu + 4 * kb + fc);
#pragma unroll
                for (int nb = 0; nb < 2; ++nb) {
                    double bf = ZC[(4 * kb + fc) * kPZ16 + 8 * nb + fr];
                    if (!MU_ZERO) bf = bf - muk;
                    dmma(acc[mb][nb][0], acc[mb][nb][1], aq[t], bf);
                }
                if (PS2(t, 2) == 1) {
                    /* the pair is complete: hand its rows of W to the chain lanes in ascending row order */
                    const int pa = mb / 2;
#pragma unroll
                    for (int u = 0; u < 2; ++u) {
                        const int mq = 2 * pa + u;
                        if (mq < NB) {
#pragma unroll
                            for (int nb = 0; nb < 2; ++nb)
                                *reinterpret_cast<double2*>(WB + fr * kPW16 + 8 * nb + 2 * fc) = make_double2(acc[mq][nb][0], acc[mq][nb][1]);
                            __syncwarp();
#pragma unroll
                            for (int r = 0; r < 8; ++r) {
                                const double w = WB[r * kPW16 + cl];
                                q = (mq == 0 && r == 0) ? w * w : fma(w, w, q);
                            }
                            __syncwarp();
                        }
                    }
                }
            }
        }
        const double lp_c = fma(-0.5, q, a.c0);
        const double loga = (lp_c - lp) + 0.0;
        if (active && -e < loga) {
#pragma unroll
            for (int i = 0; i < HR; ++i) X[(long long)(HR * half + i) * pitch + ch] = ZC[(HR * half + i) * kPZ16 + cl];
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
#pragma unroll 4
        for (int ii = 0; ii < HR; ++ii) {
            const int i = HR * half + ii;
            if (PAD && i >= a.d_real) break;                       /* padding rows are not part of the sample */
            const long long o = (long long)i * pitch + ch;
            const double v = X[o];
            if (a.sv.out) a.sv.out[(long long)i * a.sv.out_pitch + ch] = v;
            if (a.sv.sum) {
                save_moments(a.sv, o, v);
            }
        }
    }
    if (half == 0) {
        a.st.lp[ch] = lp;
        a.st.nacc[ch] = a.st.nacc[ch] + (unsigned long long)nacc;
        a.st.acc[ch] = accepted;
        if (a.sv.out) a.sv.out[(long long)(PAD ? a.d_real : D) * a.sv.out_pitch + ch] = lp;
        if (a.sv.acc_out) a.sv.acc_out[ch] = accepted;
    }
}

#undef PS1
#undef PS2

/* A fragments of a packed lower-triangular factor: frag[tile(mb,kb)][lane] = M[8mb + lane/4][4kb + lane%4] */
/* dpad >= d: the factor padded with zero rows / columns to a multiple of 8 (PAD kernels) */
static void build_frags(const double* tri_packed, int d, std::vector<double>& out, int dpad = 0) {
    const int nb = (dpad ? dpad : d) / 8;
    out.assign((size_t)nb * (nb + 1) * 32, 0.0);
    for (int mb = 0; mb < nb; ++mb)
        for (int kb = 0; kb <= 2 * mb + 1; ++kb)
            for (int lane = 0; lane < 32; ++lane) {
                const int row = 8 * mb + lane / 4, col = 4 * kb + lane % 4;
                if (col <= row && row < d) out[((size_t)(mb * (mb + 1) + kb)) * 32 + lane] = tri_packed[tri_h(row, col)];
            }
}

#ifndef AMH_MHTC_EXTRA_TU

bool mh_tc_eligible(const amh_run& r) {
    const amh_sampler& s = *r.sampler;
    const int d = r.dim;
    if (r.target->kind != AMH_TARGET_MVNORMAL) return false;
    if (!(d == 8 || d == 16 || d == 24 || d == 32)) return false;
    if (s.has_mean || s.by_components()) return false;
    if (s.d.cov_kind != AMH_COV_FULL && s.d.cov_kind != AMH_COV_DIAG && s.d.cov_kind != AMH_COV_SCALAR) return false;
    if (s.d.cov_kind != AMH_COV_FULL && r.mh_path == 2) return false;          /* the 32-chains-per-warp variant is full-covariance only */
    if (s.d.kind == AMH_SAMPLER_STATIC && !s.d.symmetric) return false;      /* needs logq: generic path */
    if (r.pitch % 32) return false;
    return true;
}

template <int D>
static int launch_mh_tc_t(amh_run& r, int nsteps, const SaveArgs& sv) {
    constexpr int WARPS = 2;
    const amh_sampler& s = *r.sampler;
    const amh_target& t = *r.target;
    if (!r.scratch) {
        std::vector<double> lf, uf, all;
        const bool covd0 = s.d.cov_kind != AMH_COV_FULL;
        if (covd0) lf.assign((size_t)(D / 8) * (D / 8 + 1) * 32, 0.0);
        else build_frags(s.scale.data(), D, lf);
        build_frags(t.blob.data() + 1 + D, D, uf);
        all = lf;
        all.insert(all.end(), uf.begin(), uf.end());
        all.insert(all.end(), t.blob.begin() + 1, t.blob.begin() + 1 + D);
        for (int i = 0; i < D; ++i)                                             /* standard deviations of a diagonal proposal */
            all.push_back(s.d.cov_kind == AMH_COV_DIAG ? s.scale[i] : s.d.cov_kind == AMH_COV_SCALAR ? s.scale[0] : 0.0);
        { const int rca = dmalloc(r.ctx, &r.scratch, all.size() * sizeof(double)); if (rca) return rca; }
        AMH_CUDA_TRY(cudaMemcpyAsync(r.scratch, all.data(), all.size() * sizeof(double), cudaMemcpyHostToDevice, r.ctx->stream));
        AMH_CUDA_TRY(sync_stream(r.ctx, r.ctx->stream));        /* `all` is a stack temporary */
    }
    constexpr int NT = (D / 8) * (D / 8 + 1);
    MhTcArgs a;
    std::memset(&a, 0, sizeof(a));
    a.st = chain_state(r);
    a.sv = sv;
    a.nsteps = nsteps;
    a.is_rw = s.d.kind == AMH_SAMPLER_RW;
    a.mu_zero = 1;
    for (int i = 0; i < D; ++i)
        if (t.blob[1 + i] != 0.0) a.mu_zero = 0;
    a.step0 = (unsigned long long)r.step;
    a.Lf = (const double*)r.scratch;
    a.Uf = a.Lf + (size_t)NT * 32;
    a.mu = a.Uf + (size_t)NT * 32;
    a.c0 = t.blob[0];
    a.dscale = a.mu + D;
    const bool covd = s.d.cov_kind != AMH_COV_FULL;
    {
        static const int pace_env = std::getenv("AMH_TC_PACE") ? std::atoi(std::getenv("AMH_TC_PACE")) : 0;
        a.pace = pace_env;
    }
    const bool v2 = r.cv == AMH_CONTRACT_V2;
    if (r.mh_path != 2 || v2) {
        /* K1T16: 16 chains per warp, 28 resident warps per SM: as ONE CTA per SM (default), or as 7 CTAs of 4 warps
         * (AMH_TC_WARPS=4, kept for A/B measurements; contract v1 only, like the 32-chain kernel K1T) */
        static const int w16_env = std::getenv("AMH_TC_WARPS") ? std::atoi(std::getenv("AMH_TC_WARPS")) : 28;
        if (w16_env == 28 || covd || v2) {
            constexpr int W28 = 28;
            const size_t smem28 = (size_t)W28 * tc16_smem_doubles_per_warp<D>() * sizeof(double);
            const unsigned grid28 = (unsigned)((r.n + 16 * W28 - 1) / (16 * W28));
            const void* key28 = (const void*)mh_step_tc16_kernel<D, W28, true, true>;
            if (!r.ctx->configured.count(key28)) {
                const char* cv = std::getenv("AMH_TC_CARVEOUT");
#define AMH_TC28_ATTR(...)                                                                                                                     \
                do {                                                                                                                               \
                    AMH_CUDA_TRY(cudaFuncSetAttribute(mh_step_tc16_kernel<D, W28, __VA_ARGS__>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem28)); \
                    if (cv) AMH_CUDA_TRY(cudaFuncSetAttribute(mh_step_tc16_kernel<D, W28, __VA_ARGS__>, cudaFuncAttributePreferredSharedMemoryCarveout, std::atoi(cv))); \
                } while (0)
                AMH_TC28_ATTR(true, true); AMH_TC28_ATTR(false, true); AMH_TC28_ATTR(true, false); AMH_TC28_ATTR(false, false);
                AMH_TC28_ATTR(true, true, true); AMH_TC28_ATTR(false, true, true); AMH_TC28_ATTR(true, false, true); AMH_TC28_ATTR(false, false, true);
                AMH_TC28_ATTR(true, true, false, 2); AMH_TC28_ATTR(false, true, false, 2); AMH_TC28_ATTR(true, false, false, 2); AMH_TC28_ATTR(false, false, false, 2);
                AMH_TC28_ATTR(true, true, true, 2); AMH_TC28_ATTR(false, true, true, 2); AMH_TC28_ATTR(true, false, true, 2); AMH_TC28_ATTR(false, false, true, 2);
#undef AMH_TC28_ATTR
                r.ctx->configured.insert(key28);
            }
#define AMH_TC28_GO(...) mh_step_tc16_kernel<D, W28, __VA_ARGS__><<<grid28, 32 * W28, smem28, r.ctx->stream>>>(a)
            if (v2) {
                if (covd) {
                    if (a.is_rw) {
                        if (a.mu_zero) AMH_TC28_GO(true, true, true, 2);
                        else AMH_TC28_GO(false, true, true, 2);
                    } else {
                        if (a.mu_zero) AMH_TC28_GO(true, false, true, 2);
                        else AMH_TC28_GO(false, false, true, 2);
                    }
                } else if (a.is_rw) {
                    if (a.mu_zero) AMH_TC28_GO(true, true, false, 2);
                    else AMH_TC28_GO(false, true, false, 2);
                } else {
                    if (a.mu_zero) AMH_TC28_GO(true, false, false, 2);
                    else AMH_TC28_GO(false, false, false, 2);
                }
            } else if (covd) {
                if (a.is_rw) {
                    if (a.mu_zero) AMH_TC28_GO(true, true, true);
                    else AMH_TC28_GO(false, true, true);
                } else {
                    if (a.mu_zero) AMH_TC28_GO(true, false, true);
                    else AMH_TC28_GO(false, false, true);
                }
            } else if (a.is_rw) {
                if (a.mu_zero) AMH_TC28_GO(true, true);
                else AMH_TC28_GO(false, true);
            } else {
                if (a.mu_zero) AMH_TC28_GO(true, false);
                else AMH_TC28_GO(false, false);
            }
#undef AMH_TC28_GO
            AMH_CUDA_TRY(cudaGetLastError());
            r.launches += 1;
            r.pending_launches += 1;
            return AMH_OK;
        }
        constexpr int W16 = 4;
        const size_t smem16 = (size_t)W16 * tc16_smem_doubles_per_warp<D>() * sizeof(double);
        const unsigned grid16 = (unsigned)((r.n + 16 * W16 - 1) / (16 * W16));
        const void* key16 = (const void*)mh_step_tc16_kernel<D, W16, true, true>;
        if (!r.ctx->configured.count(key16)) {
            /* shared memory actually needed by 7 resident CTAs; the rest of the 256 KB stays L1 for the L/U fragments */
            const char* cv = std::getenv("AMH_TC_CARVEOUT");
            const int need_kb = (int)((7 * (smem16 + 1024) + 1023) / 1024);
            const int carve = cv ? std::atoi(cv) : std::min(100, (need_kb * 100 + 227) / 228 + 1);
#define AMH_TC16_ATTR(...) AMH_CUDA_TRY(cudaFuncSetAttribute(mh_step_tc16_kernel<D, W16, __VA_ARGS__>, cudaFuncAttributePreferredSharedMemoryCarveout, carve))
            AMH_TC16_ATTR(true, true); AMH_TC16_ATTR(false, true); AMH_TC16_ATTR(true, false); AMH_TC16_ATTR(false, false);
#undef AMH_TC16_ATTR
            r.ctx->configured.insert(key16);
            if (std::getenv("AMH_TC_DEBUG")) {
                int per_sm = -1;
                cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, mh_step_tc16_kernel<D, W16, true, true>, 32 * W16, smem16);
                cudaFuncAttributes fa;
                cudaFuncGetAttributes(&fa, mh_step_tc16_kernel<D, W16, true, true>);
                std::fprintf(stderr, "[amh] K1T16 D=%d: smem/CTA %zu B, carveout %d%%, regs %d, occupancy API: %d CTAs/SM, grid %u\n", D, smem16,
                             carve, fa.numRegs, per_sm, grid16);
            }
        }
#define AMH_TC16_GO(...) mh_step_tc16_kernel<D, W16, __VA_ARGS__><<<grid16, 32 * W16, smem16, r.ctx->stream>>>(a)
        if (a.is_rw) {
            if (a.mu_zero) AMH_TC16_GO(true, true);
            else AMH_TC16_GO(false, true);
        } else {
            if (a.mu_zero) AMH_TC16_GO(true, false);
            else AMH_TC16_GO(false, false);
        }
#undef AMH_TC16_GO
        AMH_CUDA_TRY(cudaGetLastError());
        r.launches += 1;
        r.pending_launches += 1;
        return AMH_OK;
    }
    const size_t smem = (size_t)WARPS * tc_smem_doubles_per_warp<D>() * sizeof(double);
    auto kern = mh_step_tc_kernel<D, WARPS>;
    if (!r.ctx->configured.count((const void*)kern)) {
        AMH_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        AMH_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        r.ctx->configured.insert((const void*)kern);
    }
    const unsigned grid = (unsigned)((r.n + 32 * WARPS - 1) / (32 * WARPS));
    kern<<<grid, 32 * WARPS, smem, r.ctx->stream>>>(a);
    AMH_CUDA_TRY(cudaGetLastError());
    r.launches += 1;
    r.pending_launches += 1;
    return AMH_OK;
}

int launch_mh_tc(amh_run& r, int nsteps, const SaveArgs& sv) {
    switch (r.dim) {
    case 8: return launch_mh_tc_t<8>(r, nsteps, sv);
    case 16: return launch_mh_tc_t<16>(r, nsteps, sv);
    case 24: return launch_mh_tc_t<24>(r, nsteps, sv);
    case 32: return launch_mh_tc_t<32>(r, nsteps, sv);
    }
    return fail(AMH_ERR_INVALID, "tensor-core MH path: unsupported dimension");
}
#endif  /* AMH_MHTC_EXTRA_TU */

}  // namespace amhh
