/* amh_launch_ram_warp.cu -- K4W: Robust Adaptive Metropolis with ONE WARP PER CHAIN and the chain's Cholesky
 * factor resident in shared memory (RobustAdaptiveMetropolis.jl:123-173 inner step + adaptation, :216-278 steps;
 * LinearAlgebra lowrankupdate/lowrankdowndate).  Used for MvNormal targets with dim >= 16;
 * everything else runs the one-thread-per-chain kernel of amh_launch_ram.cu.
 *
 * Why: the factor is d(d+1)/2 doubles PER CHAIN (16.6 KB at d = 64, 545 MB for 32 768 chains) and a warm-up step
 * touches every element twice (S U, then the Givens sweep).  With one thread per chain the three d-vectors of a
 * step pin shared memory at ~4 warps/SM and every element access is a dependent global load: 5.6 % of the HBM
 * roofline (round-1 measurement).  Here
 *   - S lives per chain CONTIGUOUSLY, packed by COLUMNS (column i = rows i..d-1): both passes walk it column by
 *     column, lanes = rows, so every shared-memory access is one conflict-free 256-byte row;
 *   - the warp pulls its chain's factor into shared memory with ONE bulk async copy (cp.async.bulk + mbarrier, the
 *     TMA engine; 16.6 KB contiguous), runs all fused steps of the launch on it in place, and writes it back with
 *     a bulk store after every successful adaptation (that store is also the rollback point: a rejected update --
 *     eigenvalue guard :259-264, failed downdate = the reference's PosDefException -- reloads the last good factor);
 *   - HBM traffic per warm-up step = one 16.6 KB store (+ one load per chain per launch) <= the algorithmic
 *     2 x 16.6 KB; the real bound is the serial rotation recurrence (sqrt + 2 divisions per column).
 * All arithmetic is the oracle's, operation for operation (row dot products accumulate over columns in ascending
 * order, which is exactly the order of the column sweep).
 */
#include <cstdlib>
#include "amh_params.cuh"
#include "amh_fastmath.cuh"
#include "amh_ram_common.cuh"

namespace amhh {
using namespace amhd;

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned phase) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE;\n"
        "bra WAIT_LOOP;\n"
        "DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(phase) : "memory");
}
/* global -> shared bulk copy (TMA engine), completion signalled on the mbarrier */
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
/* shared -> global bulk copy */
__device__ __forceinline__ void bulk_s2g(void* dst, const void* src, unsigned bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(src)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

template <int RPL>
__host__ __device__ constexpr int ramw_doubles_per_warp(int d) {
    /* S tile (padded to 16 bytes) + U, V vectors */
    return ((d * (d + 1) / 2 + 1) & ~1) + 2 * 32 * RPL + 2;
}

/* ---------------------------------------------------------------------------------------------------------------
 * Rank-1 Cholesky up/down-date of the warp's tile (LinearAlgebra lowrankupdate / lowrankdowndate).
 * Tile layout: column i holds rows i..d-1 contiguously; lanes = rows (row j = lane + 32 r), v[r] in registers.
 *
 * The sweep is ONE serial recurrence over the columns (g_i = v_i after i rotations), ~20 dependent fp64 operations
 * per column, and only ~12 chains fit in an SM's shared memory: every instruction of the loop body is on the
 * critical path of its warp.  The fast variants are therefore
 *   - branch free: sqrt and the divisions are the correctly rounded straight-line sequences of amh_fastmath.cuh
 *     (the two quotients of a rotation share one reciprocal, a downdate column shares 1/c); operands outside their
 *     exponent range only clear the returned flag and the caller redoes the sweep with givens_sweep_slow;
 *   - split by row block (the compile-time `rb` loop), so that v[] is never indexed dynamically (it stays in
 *     registers) and rows above the diagonal block cost no predicate.
 * Results are bit-identical to the operators (tools/ubench/fdiv_probe.cu; RAM parity tests). */
template <int RPL, bool CHECK>
__device__ __forceinline__ bool givens_update_fast(double* __restrict__ Sb, int d, int lane, double (&v)[RPL],
                                                   double lo, double hi, bool& out_of_bounds) {
    int hmin = 0x7fffffff, hmax = 0;
    double* col = Sb;                                   /* col[j] = S[j][i] */
#pragma unroll
    for (int rb = 0; rb < RPL; ++rb) {
        const int iend = (32 * (rb + 1) < d) ? 32 * (rb + 1) : d;
#pragma unroll 2
        for (int i = 32 * rb; i < iend; ++i) {
            const double f = col[i];
            const double g = __shfl_sync(0xffffffffu, v[rb], i & 31);
            const double t = fma(f, f, g * g);
            hmin = min(hmin, min(__double2hiint(f), hi_abs(g)));     /* f > 0 and |g| not tiny: t >= f^2 is in range */
            hmax = max(hmax, __double2hiint(t));
            const double rr = sqrt_fast(t);
            double c, sn;
            div2_same_den(f, g, rr, c, sn);
            if (CHECK && !(lo <= rr && rr <= hi)) out_of_bounds = true;
            __syncwarp();
            if (lane == (i & 31)) col[i] = rr;
#pragma unroll
            for (int r = rb; r < RPL; ++r) {
                const int j = lane + 32 * r;
                if ((r > rb || j > i) && j < d) {
                    const double Aji = col[j];
                    const double vj = v[r];
                    col[j] = c * Aji + sn * vj;
                    v[r] = c * vj - sn * Aji;
                }
            }
            col += d - i - 1;
        }
    }
    /* f, |g| in [2^-163, ..): f^2 + g^2 >= 2^-326; t <= 2^332: rr, c, sn and every intermediate stay normal */
    return hmin >= 0x35C00000 && hmax < kHiHi;
}

template <int RPL, bool CHECK>
__device__ __forceinline__ bool givens_downdate_fast(double* __restrict__ Sb, int d, int lane, double (&v)[RPL],
                                                     double lo, double hi, bool& posdef_fail, bool& out_of_bounds) {
    int hmin = 0x7fffffff, hmax = 0;
    double* col = Sb;
    double Aii = col[0];
    double ra = rcp_refined(Aii);                      /* 1/A_ii does not depend on the recurrence: one column ahead */
#pragma unroll
    for (int rb = 0; rb < RPL; ++rb) {
        const int iend = (32 * (rb + 1) < d) ? 32 * (rb + 1) : d;
#pragma unroll 2
        for (int i = 32 * rb; i < iend; ++i) {
            const double g = __shfl_sync(0xffffffffu, v[rb], i & 31);
            double* coln = col + (d - i - 1);
            const double Ann = (i + 1 < d) ? coln[i + 1] : 1.0;
            const double sn = div_with_rcp(g, Aii, ra);
            const double s2 = sn * sn;
            if (s2 > 1.0) posdef_fail = true;          /* the reference throws here; the tile is rolled back */
            const double om = 1.0 - s2;
            if (!posdef_fail) {                        /* (after a failure nothing of the sweep is kept) */
                hmin = min(hmin, min(min(__double2hiint(Aii), hi_abs(g)), __double2hiint(om)));
                hmax = max(hmax, max(__double2hiint(Aii), hi_abs(g)));
            }
            const double c = sqrt_fast(om);
            const double rc = rcp_refined(c);
            const double dg = c * Aii;
            if (CHECK && !posdef_fail && !(lo <= dg && dg <= hi)) out_of_bounds = true;
            __syncwarp();
            if (lane == (i & 31)) col[i] = dg;
#pragma unroll
            for (int r = rb; r < RPL; ++r) {
                const int j = lane + 32 * r;
                if ((r > rb || j > i) && j < d) {
                    const double num = col[j] - sn * v[r];
                    if (!posdef_fail) {
                        hmin = min(hmin, hi_abs(num));
                        hmax = max(hmax, hi_abs(num));
                    }
                    const double Aji = div_with_rcp(num, c, rc);
                    col[j] = Aji;
                    v[r] = -sn * Aji + c * v[r];
                }
            }
            Aii = Ann;
            ra = rcp_refined(Ann);
            col = coln;
        }
    }
    return hmin >= kHiLo && hmax < kHiHi;
}

/* the same sweeps with the IEEE operators (any operands); out of line: it is the redo path */
template <int RPL>
__device__ __noinline__ void givens_sweep_slow(double* __restrict__ Sb, int d, int lane, double (&v)[RPL], bool update, bool check,
                                               double lo, double hi, bool& posdef_fail, bool& out_of_bounds) {
    for (int i = 0; i < d; ++i) {
        double* col = Sb + colstart(i, d) - i;
        const double f = col[i];
        double g = 0.0;
#pragma unroll
        for (int r = 0; r < RPL; ++r)
            if ((i >> 5) == r) g = __shfl_sync(0xffffffffu, v[r], i & 31);
        if (update) {
            const double rr = sqrt(fma(f, f, g * g));
            const double c = f / rr, sn = g / rr;
            if (check && !(lo <= rr && rr <= hi)) out_of_bounds = true;
            __syncwarp();
            if (lane == (i & 31)) col[i] = rr;
#pragma unroll
            for (int r = 0; r < RPL; ++r) {
                const int j = lane + 32 * r;
                if (j > i && j < d) {
                    const double Aji = col[j];
                    const double vj = v[r];
                    col[j] = c * Aji + sn * vj;
                    v[r] = c * vj - sn * Aji;
                }
            }
        } else {
            const double sn = g / f;
            const double s2 = sn * sn;
            if (s2 > 1.0) { posdef_fail = true; break; }
            const double c = sqrt(1.0 - s2);
            const double dg = c * f;
            if (check && !(lo <= dg && dg <= hi)) out_of_bounds = true;
            __syncwarp();
            if (lane == (i & 31)) col[i] = dg;
#pragma unroll
            for (int r = 0; r < RPL; ++r) {
                const int j = lane + 32 * r;
                if (j > i && j < d) {
                    const double Aji = (col[j] - sn * v[r]) / c;
                    col[j] = Aji;
                    v[r] = -sn * Aji + c * v[r];
                }
            }
        }
    }
}

/* DF = the dimension as a compile-time constant (0: runtime).  With DF the column loops unroll completely: column
 * offsets and row predicates of the two mat-vecs and the two sums become immediates. */
template <int RPL, int DF>
__global__ void __launch_bounds__(512)
ram_warp_kernel(const __grid_constant__ RamWArgs a) {
    constexpr int UNRC = DF ? DF : 4;          /* matvec / sum loops */
    extern __shared__ __align__(16) double smem_w[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int WARPS = blockDim.x >> 5;
    const int d = DF ? DF : a.d;
    const int nt = d * (d + 1) / 2;
    const int ntp = (nt + 1) & ~1;
    /* the target factor (column-packed) is shared by the CTA's warps */
    double* __restrict__ Ut = smem_w;
    for (int i = threadIdx.x; i < nt; i += blockDim.x) Ut[i] = a.Utc[i];
    __syncthreads();
    double* __restrict__ Sb = smem_w + ntp + (size_t)warp * ramw_doubles_per_warp<RPL>(d);
    double* __restrict__ Us = Sb + ntp;                  /* U (noise), [32*RPL] */
    double* __restrict__ Vs = Us + 32 * RPL;             /* x_new - mu, then w  */
    unsigned long long* bar = reinterpret_cast<unsigned long long*>(Vs + 32 * RPL);
    const unsigned sbytes = (unsigned)(ntp * sizeof(double));   /* bulk copies move multiples of 16 bytes */
    const long long pitch = a.st.pitch;
    const int cv = a.st.cv;
    const unsigned long long B = amh::blocks_per_step_cv(cv, d);
    const int npb = (d + 1) / 2;                         /* normal PAIRS of a step */
    const int nbe = amh::normal_blocks(cv, d);           /* block of the exponential */
    if (lane == 0) mbar_init(bar, 1);
    __syncwarp();
    unsigned phase = 0;
    const long long nwarps = (long long)gridDim.x * WARPS;

    for (long long ch = (long long)blockIdx.x * WARPS + warp; ch < a.st.n; ch += nwarps) {
        double* __restrict__ Sg = a.S + (size_t)ch * ntp;
        /* pull the factor: one bulk copy per chain and launch */
        if (lane == 0) {
            mbar_expect_tx(bar, sbytes);
            bulk_g2s(Sb, Sg, sbytes, bar);
        }
        const unsigned long long seed = a.st.seeds[ch];
        double lp = a.st.lp[ch], logalpha = a.logalpha[ch], eta = a.eta[ch];
        unsigned nacc = 0u;
        unsigned char accepted = a.st.acc[ch], failed = a.failed[ch];
        double x[RPL];
#pragma unroll
        for (int r = 0; r < RPL; ++r) {
            const int j = lane + 32 * r;
            x[r] = (j < d) ? a.st.X[(long long)j * pitch + ch] : 0.0;
        }
        mbar_wait(bar, phase);
        phase ^= 1u;

        for (int s = 0; s < a.nsteps; ++s) {
            const unsigned long long k = a.step0 + (unsigned long long)s + 1ull;   /* = state.iteration */
            const unsigned long long blk0 = k * B;
            /* U = randn(rng, d)  (:135): lane l draws blocks l, l+32, ...; every lane draws the exponential */
            if (cv == AMH_CONTRACT_V2) {
                /* contract v2: pair p = normals 2p, 2p+1 comes from words (2(p&1), 2(p&1)+1) of block p >> 1; two lanes
                 * run the same 7-round block and keep one pair each (shorter critical path than one lane doing both) */
                for (int p = lane; p < npb; p += 32) {
                    const amh::Block b = amh::stream_block7(seed, blk0 + (unsigned long long)(p >> 1), 0u);
                    double z0, z1;
                    amh::normal_pair32((p & 1) ? b.v[2] : b.v[0], (p & 1) ? b.v[3] : b.v[1], z0, z1);
                    Us[2 * p] = z0;
                    if (2 * p + 1 < d) Us[2 * p + 1] = z1;
                }
            } else {
                for (int jb = lane; jb < npb; jb += 32) {
                    const amh::Block b = amh::stream_block(seed, blk0 + (unsigned long long)jb, 0u);
                    double z0, z1;
                    amh::normal_pair(b, z0, z1);
                    Us[2 * jb] = z0;
                    if (2 * jb + 1 < d) Us[2 * jb + 1] = z1;
                }
            }
            const amh::Block be = amh::step_block(cv, seed, blk0 + (unsigned long long)nbe);
            const double e = amh::exponential(be.v[0], be.v[1]);
            __syncwarp();
            /* y = S U by columns; x_new = muladd(S, U, x)  (:136) */
            double y[RPL], xn[RPL];
#pragma unroll
            for (int r = 0; r < RPL; ++r) y[r] = 0.0;
#pragma unroll UNRC
            for (int i = 0; i < d; ++i) {
                const double ui = Us[i];
                const double* col = Sb + colstart(i, d) - i;
#pragma unroll
                for (int r = 0; r < RPL; ++r) {
                    const int j = lane + 32 * r;
                    if (j >= i && j < d) y[r] = (i == 0) ? col[j] * ui : fma(col[j], ui, y[r]);
                }
            }
#pragma unroll
            for (int r = 0; r < RPL; ++r) {
                const int j = lane + 32 * r;
                xn[r] = y[r] + x[r];
                if (j < d) Vs[j] = xn[r] - __ldg(a.mu + j);
            }
            __syncwarp();
            /* target: w = U_t (x_new - mu) by columns, q = sum w_j^2 in index order */
            double w[RPL];
#pragma unroll
            for (int r = 0; r < RPL; ++r) w[r] = 0.0;
#pragma unroll UNRC
            for (int i = 0; i < d; ++i) {
                const double vi = Vs[i];
                const double* col = Ut + colstart(i, d) - i;
#pragma unroll
                for (int r = 0; r < RPL; ++r) {
                    const int j = lane + 32 * r;
                    if (j >= i && j < d) w[r] = (i == 0) ? col[j] * vi : fma(col[j], vi, w[r]);
                }
            }
            __syncwarp();
#pragma unroll
            for (int r = 0; r < RPL; ++r) {
                const int j = lane + 32 * r;
                if (j < d) Vs[j] = w[r];
            }
            __syncwarp();
            double q = Vs[0] * Vs[0];
#pragma unroll UNRC
            for (int j = 1; j < d; ++j) q = fma(Vs[j], Vs[j], q);
            const double lp_new = fma(-0.5, q, a.c0);
            const double dl = lp_new - lp;
            logalpha = (dl != dl) ? dl : (dl < 0.0 ? dl : 0.0);        /* min(lp_new - lp, 0)  (:147) */
            const bool isaccept = e > -logalpha;                       /* (:148) */
            if (a.warmup) {
                /* ram_adapt (:153-173) */
                const double dalpha = amh::exp_(logalpha) - a.alpha;
                eta = amh::exp_(-a.gamma * amh::log_((double)(long long)k));
                if (dalpha == dalpha) {
                    const double cc = sqrt(eta * fabs(dalpha));
                    double nu = Us[0] * Us[0];
#pragma unroll UNRC
                    for (int i = 1; i < d; ++i) nu = fma(Us[i], Us[i], nu);
                    nu = sqrt(nu);
                    double v[RPL];
                    const bool nfast = nu > 1e-100 && nu < 1e100;
                    const double rnu = rcp_refined(nu);
#pragma unroll
                    for (int r = 0; r < RPL; ++r) {
                        const double num = cc * y[r];
                        v[r] = div_with_rcp(num, nu, rnu);
                        if (!(nfast && fabs(num) > 1e-100 && fabs(num) < 1e100)) v[r] = div_slow(num, nu);
                    }
                    if (lane == 0) bulk_wait_read();  /* the previous write-back has finished reading the tile */
                    __syncwarp();
                    bool posdef_fail = false, out_of_bounds = false;
                    /* speculative branch-free sweep; `ok` is false when some operand left the exponent range the
                     * straight-line sqrt / division sequences cover (never, in practice) */
                    bool ok;
                    if (a.check)
                        ok = (dalpha > 0.0) ? givens_update_fast<RPL, true>(Sb, d, lane, v, a.lo, a.hi, out_of_bounds)
                                            : givens_downdate_fast<RPL, true>(Sb, d, lane, v, a.lo, a.hi, posdef_fail, out_of_bounds);
                    else
                        ok = (dalpha > 0.0) ? givens_update_fast<RPL, false>(Sb, d, lane, v, a.lo, a.hi, out_of_bounds)
                                            : givens_downdate_fast<RPL, false>(Sb, d, lane, v, a.lo, a.hi, posdef_fail, out_of_bounds);
                    ok = __all_sync(0xffffffffu, ok) && !a.force_redo;
                    if (!ok) {
                        /* redo with the IEEE operators from the last good factor (global memory holds it) */
                        if (lane == 0) {
                            fence_async_smem();
                            bulk_wait_all();
                            mbar_expect_tx(bar, sbytes);
                            bulk_g2s(Sb, Sg, sbytes, bar);
                        }
                        mbar_wait(bar, phase);
                        phase ^= 1u;
#pragma unroll
                        for (int r = 0; r < RPL; ++r) v[r] = div_slow(cc * y[r], nu);
                        posdef_fail = false;
                        out_of_bounds = false;
                        givens_sweep_slow<RPL>(Sb, d, lane, v, dalpha > 0.0, a.check != 0, a.lo, a.hi, posdef_fail, out_of_bounds);
                    }
                    if (posdef_fail) failed = 1;
                    __syncwarp();
                    if (posdef_fail || out_of_bounds) {
                        /* S is kept (:259-264): roll the tile back to the last good factor */
                        if (lane == 0) {
                            fence_async_smem();
                            bulk_wait_all();
                            mbar_expect_tx(bar, sbytes);
                            bulk_g2s(Sb, Sg, sbytes, bar);
                        }
                        mbar_wait(bar, phase);
                        phase ^= 1u;
                    } else {
                        /* S_new becomes current: write it through (also the next rollback point) */
                        fence_async_smem();
                        __syncwarp();
                        if (lane == 0) bulk_s2g(Sg, Sb, sbytes);
                    }
                } else {
                    failed = 1;
                }
            }
            if (isaccept) {
#pragma unroll
                for (int r = 0; r < RPL; ++r) x[r] = xn[r];
                lp = lp_new;
                ++nacc;
            }
            accepted = isaccept ? 1 : 0;
            __syncwarp();
        }

        /* write the chain state back */
#pragma unroll
        for (int r = 0; r < RPL; ++r) {
            const int j = lane + 32 * r;
            if (j < d) {
                const long long o = (long long)j * pitch + ch;
                a.st.X[o] = x[r];
                if (a.sv.out) a.sv.out[(long long)j * a.sv.out_pitch + ch] = x[r];
                if (a.sv.sum) {
                    save_moments(a.sv, o, x[r]);
                }
            }
        }
        if (lane == 0) {
            a.st.lp[ch] = lp;
            a.st.nacc[ch] = a.st.nacc[ch] + (unsigned long long)nacc;
            a.st.acc[ch] = accepted;
            a.logalpha[ch] = logalpha;
            a.eta[ch] = eta;
            a.failed[ch] = failed;
            if (a.sv.out) a.sv.out[(long long)d * a.sv.out_pitch + ch] = lp;
            if (a.sv.acc_out) a.sv.acc_out[ch] = accepted;
            bulk_wait_all();                          /* the tile is about to be overwritten by the next chain */
        }
        __syncwarp();
    }
}

/* ---- layout conversion: [chain][column-packed] <-> the ABI's [row-packed tri][chain] ---- */
__global__ void ramw_export_S_kernel(const double* S1, const double* S2, const unsigned char* sflag, double* dst, long long n, long long pitch, int d) {
    const long long ch = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (ch >= n) return;
    const int ntp = (d * (d + 1) / 2 + 1) & ~1;
    const double* Sw = (sflag && sflag[ch]) ? S2 : S1;          /* K4S: the chain's current buffer */
    for (int i = 0; i < d; ++i)
        for (int j = 0; j <= i; ++j) dst[(long long)tri(i, j) * pitch + ch] = Sw[(size_t)ch * ntp + colstart(j, d) + (i - j)];
}
__global__ void ramw_import_S_kernel(double* S1, double* S2, const unsigned char* sflag, const double* src, long long n, long long pitch, int d) {
    const long long ch = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (ch >= n) return;
    const int ntp = (d * (d + 1) / 2 + 1) & ~1;
    double* Sw = (sflag && sflag[ch]) ? S2 : S1;
    for (int i = 0; i < d; ++i)
        for (int j = 0; j <= i; ++j) Sw[(size_t)ch * ntp + colstart(j, d) + (i - j)] = src[(long long)tri(i, j) * pitch + ch];
}
__global__ void ramw_init_S_kernel(double* Sw, const double* S0 /* row-packed or NULL */, long long n, int d) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int nt = d * (d + 1) / 2;
    const int ntp = (nt + 1) & ~1;
    if (idx >= n * ntp) return;
    const int q = (int)(idx % ntp);
    if (q >= nt) { Sw[idx] = 0.0; return; }
    /* column-packed position q -> (row, col) */
    int col = 0, rem = q;
    while (rem >= d - col) { rem -= d - col; ++col; }
    const int row = col + rem;
    Sw[idx] = S0 ? S0[tri(row, col)] : (row == col ? 1.0 : 0.0);
}

bool ram_warp_eligible(const amh_run& r) {
    return r.sampler->d.kind == AMH_SAMPLER_RAM && r.target->kind == AMH_TARGET_MVNORMAL && r.dim >= 16 && r.dim <= 128;
}

int ramw_init_S(amh_run& r) {
    const long long total = r.n * (long long)((r.dim * (r.dim + 1) / 2 + 1) & ~1);
    const unsigned grid = (unsigned)((total + 255) / 256);
    ramw_init_S_kernel<<<grid, 256, 0, r.ctx->stream>>>(r.S, r.sampler->dS0, r.n, r.dim);
    AMH_CUDA_TRY(cudaGetLastError());
    r.launches += 1;
    return AMH_OK;
}

int ramw_export_S(amh_run& r, double* dst) {
    const unsigned grid = (unsigned)((r.n + 127) / 128);
    ramw_export_S_kernel<<<grid, 128, 0, r.ctx->stream>>>(r.S, r.S2, r.sflag, dst, r.n, r.pitch, r.dim);
    AMH_CUDA_TRY(cudaGetLastError());
    r.launches += 1;
    return AMH_OK;
}

int ramw_import_S(amh_run& r, const double* src) {
    const unsigned grid = (unsigned)((r.n + 127) / 128);
    ramw_import_S_kernel<<<grid, 128, 0, r.ctx->stream>>>(r.S, r.S2, r.sflag, src, r.n, r.pitch, r.dim);
    AMH_CUDA_TRY(cudaGetLastError());
    r.launches += 1;
    return AMH_OK;
}

template <int RPL, int DF>
static int launch_ram_warp_t(amh_run& r, int nsteps, bool warmup, const SaveArgs& sv) {
    const amh_sampler& s = *r.sampler;
    const amh_target& t = *r.target;
    const int d = r.dim;
    const int nt = d * (d + 1) / 2;
    const int ntp = (nt + 1) & ~1;
    if (!r.scratch) {
        /* target factor re-packed by columns + mu */
        std::vector<double> all((size_t)nt + d);
        const double* U = t.blob.data() + 1 + d;
        for (int col = 0; col < d; ++col)
            for (int row = col; row < d; ++row) all[colstart(col, d) + (row - col)] = U[tri_h(row, col)];
        for (int i = 0; i < d; ++i) all[(size_t)nt + i] = t.blob[1 + i];
        const int rca = dmalloc(r.ctx, &r.scratch, all.size() * sizeof(double));
        if (rca) return rca;
        AMH_CUDA_TRY(cudaMemcpyAsync(r.scratch, all.data(), all.size() * sizeof(double), cudaMemcpyHostToDevice, r.ctx->stream));
        AMH_CUDA_TRY(cudaStreamSynchronize(r.ctx->stream));
    }
    RamWArgs a;
    std::memset(&a, 0, sizeof(a));
    a.st = chain_state(r);
    a.sv = sv;
    a.d = d;
    a.nsteps = nsteps;
    a.warmup = warmup ? 1 : 0;
    a.step0 = (unsigned long long)r.step;
    a.S = r.S; a.failed = r.failed; a.logalpha = r.logalpha; a.eta = r.eta;
    a.alpha = s.d.ram_alpha; a.gamma = s.d.ram_gamma; a.lo = s.d.ram_eig_lo; a.hi = s.d.ram_eig_hi;
    a.check = !(a.lo == 0.0 && a.hi == INFINITY);
    a.Utc = (const double*)r.scratch;
    a.force_redo = std::getenv("AMH_RAMW_FORCE_REDO") != nullptr;      /* test switch, read per launch */
    a.mu = a.Utc + nt;
    a.c0 = t.blob[0];
    /* one CTA per SM holding as many chain tiles as fit beside the shared target factor */
    const size_t per_warp = (size_t)ramw_doubles_per_warp<RPL>(d) * sizeof(double);
    const size_t fixed = (size_t)ntp * sizeof(double);
    const size_t budget = 226 * 1024;
    int warps = (int)((budget - fixed) / per_warp);
    if (warps > 16) warps = 16;
    if (warps < 1) return fail(AMH_ERR_UNSUPPORTED, "RAM warp kernel: factor does not fit in shared memory");
    const size_t smem = fixed + (size_t)warps * per_warp;
    auto kern = ram_warp_kernel<RPL, DF>;
    AMH_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    AMH_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 32 * warps, smem));
    if (per_sm < 1) return fail(AMH_ERR_UNSUPPORTED, "RAM warp kernel: factor does not fit in shared memory");
    const long long want = (r.n + warps - 1) / warps;
    const long long cap = (long long)per_sm * r.ctx->sm_count;
    const unsigned grid = (unsigned)std::min<long long>(want, cap);        /* persistent: warps loop over chains */
    kern<<<grid, 32 * warps, smem, r.ctx->stream>>>(a);
    AMH_CUDA_TRY(cudaGetLastError());
    r.launches += 1;
    r.pending_launches += 1;
    return AMH_OK;
}

int launch_ram_warp(amh_run& r, int nsteps, bool warmup, const SaveArgs& sv) {
    static const bool generic_only = std::getenv("AMH_RAMW_GENERIC") != nullptr;     /* A/B switch, tests */
    if (!generic_only) {
        if (r.dim == 64) return launch_ram_warp_t<2, 64>(r, nsteps, warmup, sv);       /* BASELINE config 5 */
        if (r.dim == 32) return launch_ram_warp_t<1, 32>(r, nsteps, warmup, sv);
    }
    const int rpl = (r.dim + 31) / 32;
    switch (rpl) {
    case 1: return launch_ram_warp_t<1, 0>(r, nsteps, warmup, sv);
    case 2: return launch_ram_warp_t<2, 0>(r, nsteps, warmup, sv);
    case 3: return launch_ram_warp_t<3, 0>(r, nsteps, warmup, sv);
    case 4: return launch_ram_warp_t<4, 0>(r, nsteps, warmup, sv);
    }
    return fail(AMH_ERR_UNSUPPORTED, "RAM warp kernel supports dim <= 128");
}

}  // namespace amhh
