/* amh_launch_mala_tensor.cu -- K3T: MALA on the many-row logistic-regression target (BASELINE config 4) with the two
 * design-matrix contractions of a step on the 5th-generation tensor cores (tcgen05.mma, accumulators in TMEM, operands
 * staged by TMA), as SPLIT-bf16 GEMMs with fp32 accumulation.  OPT-IN (amh_sampler_desc.precision ==
 * AMH_PRECISION_BF16X2): unlike every other kernel of the library this one is NOT bit-exact against the oracle -- it
 * trades the last ~10 bits of the log-density for the 1.4 PFLOP/s bf16 datapath where K3L is capped by the 37 TFLOP/s
 * FP64 pipe.  The tolerance is stated and tested (tests/test_parity_baseline_gpu.py::test_mala_tensor_*).
 *
 * Reference semantics (MALA.jl:54-93 with logdensity_and_gradient of the logistic model, MALA.jl:100-105):
 *   cand = x + (sigma2/2) grad(x) + sigma z;   eta = X cand;   lp = sum_i [y_i eta_i - log1pexp(eta_i)] - |cand|^2/(2 tau^2)
 *   grad(cand) = X' (y - sigmoid(eta)) - cand / tau^2;   Hastings ratio from both drifts;   accept iff -randexp < log-alpha.
 * Everything except `eta = X cand` and `X' r` is the fp64 arithmetic of K3L / the oracle (candidate, noise, prior term,
 * Hastings terms, accept); those two are
 *   GEMM1  eta^T[chains x rows]  = C[chains x d] X^T          M = 128 chains, N = 64 rows per tile, K = d = 128
 *   GEMM2  gg^T [chains x d]     = R[chains x rows] X          M = 128 chains, N = d = 128,          K = 64 rows per tile
 * with C = C_hi + C_lo, X = X_hi + X_lo, R = R_hi + R_lo in bf16 and the three products hi*hi + hi*lo + lo*hi (relative
 * error ~2^-16 per product); the elementwise link (log1pexp, sigmoid) runs in fp32 between the two GEMMs, per chain,
 * on the warps that own the accumulator's TMEM lanes, so eta never leaves the SM.
 *
 * One CTA = 128 chains (one per TMEM lane) for one MCMC step; 10 warps:
 *   warp 0    TMA producer: per 64-row tile, X[64 x 128] (GEMM1's B) and X^T[128 x 64] (GEMM2's B), hi and lo, 64 KB,
 *             into a 2-stage ring (cp.async.bulk.tensor.2d, SWIZZLE_128B, full / empty mbarriers)
 *   warp 1    MMA issuer (one elected lane): GEMM1(t) into one of two TMEM accumulators, then GEMM2(t-1) as soon as the
 *             epilogue warps have published R(t-1); tcgen05.commit releases smem stages / signals accumulators
 *   warps 2-9 two threads per chain (halves of the dimensions / of a tile's columns): candidate (fp64, contract noise)
 *             -> C_hi / C_lo written straight into the swizzled
 *             K-major operand layout; per tile tcgen05.ld eta, link function, running log-likelihood, R_hi / R_lo into
 *             shared memory (operand A of GEMM2); at the end tcgen05.ld gg and the fp64 tail of the step.
 * Shared memory 224 KB (C 64 + ring 128 + R 32), TMEM 256 columns (2 x 64 eta + 128 gg). */
#include <cuda.h>
#include <cuda_bf16.h>
#include <cstdlib>
#include <mutex>
#include "amh_params.cuh"

namespace amhh {
using namespace amhd;

constexpr int kTD = 128;                 /* dimension this kernel is built for */
constexpr int kTRows = 64;               /* design-matrix rows per tile */

struct MalaTArgs {
    ChainState st;
    SaveArgs sv;
    unsigned long long step;             /* the step this launch takes (1-based) */
    double sigma, sigma2, drift, inv2tau2, invtau2;
    int nrows, ntiles;
    const float* y;                      /* [ntiles * 64], zero beyond nrows */
    double* Xc;                          /* [D][pitch] candidate */
    double* Gc;                          /* [D][pitch] gradient at the candidate */
};

/* ---- PTX wrappers (checked in isolation by tools/ubench/umma_probe.cu) ---- */
__device__ __forceinline__ unsigned t_smem(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void t_mbar_init(unsigned long long* b, int c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(t_smem(b)), "r"(c)); }
__device__ __forceinline__ void t_mbar_wait(unsigned long long* b, unsigned ph) {
    asm volatile("{\n.reg .pred p;\nTW:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra TD;\nbra TW;\nTD:\n}\n" ::"r"(t_smem(b)), "r"(ph) : "memory");
}
__device__ __forceinline__ void t_mbar_arrive(unsigned long long* b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(t_smem(b)) : "memory"); }
__device__ __forceinline__ void t_mbar_expect_tx(unsigned long long* b, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(t_smem(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void t_tma_2d(void* dst, const CUtensorMap* map, int c0, int c1, unsigned long long* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(t_smem(dst)), "l"(map), "r"(c0), "r"(c1), "r"(t_smem(bar)) : "memory");
}
/* K-major SWIZZLE_128B shared-memory descriptor: 128-byte rows, 8-row groups 1024 bytes apart, descriptor version 1 */
__device__ __forceinline__ unsigned long long t_desc(const void* p) {
    return (unsigned long long)((t_smem(p) >> 4) & 0x3FFF) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
/* instruction descriptor: D = f32, A = B = bf16, both K-major, N >> 3 at bit 17, M >> 4 at bit 24 */
__host__ __device__ constexpr unsigned t_idesc(int M, int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((unsigned)(N >> 3) << 17) | ((unsigned)(M >> 4) << 24);
}
__device__ __forceinline__ void t_umma(unsigned tmem_d, unsigned long long da, unsigned long long db, unsigned idesc, unsigned acc) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n"
                 ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void t_commit(unsigned long long* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(t_smem(bar)) : "memory");
}
__device__ __forceinline__ void t_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void t_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void t_fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void t_ld16(unsigned taddr, float (&v)[16]) {
    unsigned r[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                   "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

/* MUFU approximations without the denormal fix-ups of __expf / __logf (arguments here are never denormal) */
__device__ __forceinline__ float ex2_approx(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float lg2_approx(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcp_approx(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
/* 8 consecutive K elements (one 16-byte chunk) of operand row `row` in a [rows x 64] K-block tile, SWIZZLE_128B */
__device__ __forceinline__ void store_chunk(unsigned char* tile, int row, int chunk, const unsigned (&e)[4]) {
    *reinterpret_cast<uint4*>(tile + row * 128 + ((chunk ^ (row & 7)) << 4)) = make_uint4(e[0], e[1], e[2], e[3]);
}

/* shared-memory map (bytes) */
constexpr int kOffC = 0;                          /* [slice 2][kblock 2][128 chains x 128 B]          65 536 */
constexpr int kOffX = 65536;                      /* [stage 2][slice 2][kblock 2][64 rows x 128 B]    65 536 */
constexpr int kOffXT = 131072;                    /* [stage 2][slice 2][128 features x 128 B]         65 536 */
constexpr int kOffR = 196608;                     /* [slice 2][128 chains x 128 B]                    32 768 */
constexpr int kOffBar = 229376;
constexpr int kSmemT = kOffBar + 160;
enum { BAR_FULL = 0, BAR_EMPTY = 2, BAR_A1FULL = 4, BAR_A1EMPTY = 6, BAR_RFULL = 8, BAR_REMPTY = 9, BAR_A2FULL = 10, BAR_CREADY = 11,
       BAR_FULLT = 12, BAR_EMPTYT = 14, BAR_COUNT = 16 };     /* FULL / EMPTY: the X half of a stage; FULLT / EMPTYT: the X^T half */

/* NQ = threads per chain = epilogue warps per TMEM lane group (2: 8 epilogue warps, 4: 16) */
template <int NQ>
__global__ void __launch_bounds__(64 + 128 * NQ, 1)
mala_tensor_kernel(const __grid_constant__ MalaTArgs a, const __grid_constant__ CUtensorMap mapXhi, const __grid_constant__ CUtensorMap mapXlo,
                   const __grid_constant__ CUtensorMap mapXThi, const __grid_constant__ CUtensorMap mapXTlo) {
    extern __shared__ __align__(1024) unsigned char tsm[];
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(tsm + kOffBar);
    unsigned* tmem_ptr = reinterpret_cast<unsigned*>(tsm + kOffBar + BAR_COUNT * 8);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int T = a.ntiles;
    /* every CTA walks the row tiles in its own rotation: the sums over tiles do not care about the order, and 128 CTAs
     * asking L2 for the SAME 64 KB at the same moment serialise on its slices (ncu: the epilogue warps spent 21 % of their
     * stall samples waiting for GEMM1, i.e. for the tile's TMA) */
    const int trot = (int)((37u * blockIdx.x) % (unsigned)T);

    if (threadIdx.x == 0) {
        t_mbar_init(bars + BAR_FULL, 1); t_mbar_init(bars + BAR_FULL + 1, 1);
        t_mbar_init(bars + BAR_EMPTY, 1); t_mbar_init(bars + BAR_EMPTY + 1, 1);
        t_mbar_init(bars + BAR_FULLT, 1); t_mbar_init(bars + BAR_FULLT + 1, 1);
        t_mbar_init(bars + BAR_EMPTYT, 1); t_mbar_init(bars + BAR_EMPTYT + 1, 1);
        t_mbar_init(bars + BAR_A1FULL, 1); t_mbar_init(bars + BAR_A1FULL + 1, 1);
        t_mbar_init(bars + BAR_A1EMPTY, 4 * NQ); t_mbar_init(bars + BAR_A1EMPTY + 1, 4 * NQ);
        t_mbar_init(bars + BAR_RFULL, 4 * NQ);
        t_mbar_init(bars + BAR_REMPTY, 1);
        t_mbar_init(bars + BAR_A2FULL, 1);
        t_mbar_init(bars + BAR_CREADY, 4 * NQ);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(t_smem(tmem_ptr)), "r"(256));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    t_fence_before();
    __syncthreads();
    t_fence_after();
    const unsigned tmem = *tmem_ptr;

    if (warp == 0) {
        /* ===== TMA producer ===== */
        if (lane == 0) {
            for (int t = 0; t < T; ++t) {
                const int s = t & 1, u = t >> 1;
                int tt = t + trot; if (tt >= T) tt -= T;
                /* the X half (GEMM1's B) is released as soon as GEMM1 of tile t-2 has completed, the X^T half (GEMM2's B)
                 * after GEMM2 of tile t-2: two barrier pairs per stage keep the loads off the GEMM2 -> GEMM1 chain */
                unsigned char* dx = tsm + kOffX + s * 32768;
                unsigned char* dxt = tsm + kOffXT + s * 32768;
                t_mbar_wait(bars + BAR_EMPTY + s, (unsigned)((u & 1) ^ 1));
                t_mbar_expect_tx(bars + BAR_FULL + s, 32768u);
                for (int kb = 0; kb < 2; ++kb) {
                    t_tma_2d(dx + kb * 8192, &mapXhi, kb * 64, tt * kTRows, bars + BAR_FULL + s);
                    t_tma_2d(dx + 16384 + kb * 8192, &mapXlo, kb * 64, tt * kTRows, bars + BAR_FULL + s);
                }
                t_mbar_wait(bars + BAR_EMPTYT + s, (unsigned)((u & 1) ^ 1));
                t_mbar_expect_tx(bars + BAR_FULLT + s, 32768u);
                t_tma_2d(dxt, &mapXThi, tt * kTRows, 0, bars + BAR_FULLT + s);
                t_tma_2d(dxt + 16384, &mapXTlo, tt * kTRows, 0, bars + BAR_FULLT + s);
            }
        }
    } else if (warp == 1) {
        /* ===== MMA issuer ===== */
        if (lane == 0) {
            constexpr unsigned id1 = t_idesc(128, kTRows), id2 = t_idesc(128, kTD);
            const unsigned acc2 = tmem + 128;
            auto gemm2 = [&](int tt) {
                const int s = tt & 1;
                t_mbar_wait(bars + BAR_FULLT + s, (unsigned)((tt >> 1) & 1));
                t_mbar_wait(bars + BAR_RFULL, (unsigned)(tt & 1));
                t_fence_after();
                const unsigned char* sr = tsm + kOffR;
                const unsigned char* sxt = tsm + kOffXT + s * 32768;
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) {
                    /* R_hi X_hi + R_hi X_lo + R_lo X_hi */
                    t_umma(acc2, t_desc(sr + kk * 32), t_desc(sxt + kk * 32), id2, (tt | kk) ? 1u : 0u);
                    t_umma(acc2, t_desc(sr + kk * 32), t_desc(sxt + 16384 + kk * 32), id2, 1u);
                    t_umma(acc2, t_desc(sr + 16384 + kk * 32), t_desc(sxt + kk * 32), id2, 1u);
                }
                t_commit(bars + BAR_REMPTY);            /* R may be overwritten */
                t_commit(bars + BAR_EMPTYT + s);        /* the stage's X^T half may be overwritten */
            };
            t_mbar_wait(bars + BAR_CREADY, 0u);
            t_fence_after();
            for (int t = 0; t < T; ++t) {
                const int s = t & 1, u = t >> 1;
                t_mbar_wait(bars + BAR_FULL + s, (unsigned)(u & 1));
                t_mbar_wait(bars + BAR_A1EMPTY + s, (unsigned)((u & 1) ^ 1));
                t_fence_after();
                const unsigned acc1 = tmem + (unsigned)(s * 64);
                const unsigned char* sc = tsm + kOffC;
                const unsigned char* sx = tsm + kOffX + s * 32768;
#pragma unroll
                for (int kb = 0; kb < 2; ++kb)
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk) {
                        const int oc = kb * 16384 + kk * 32, ox = kb * 8192 + kk * 32;
                        /* C_hi X_hi + C_hi X_lo + C_lo X_hi */
                        t_umma(acc1, t_desc(sc + oc), t_desc(sx + ox), id1, (kb | kk) ? 1u : 0u);
                        t_umma(acc1, t_desc(sc + oc), t_desc(sx + 16384 + ox), id1, 1u);
                        t_umma(acc1, t_desc(sc + 32768 + oc), t_desc(sx + ox), id1, 1u);
                    }
                t_commit(bars + BAR_A1FULL + s);
                t_commit(bars + BAR_EMPTY + s);         /* the stage's X half may be overwritten */
                if (t >= 1) gemm2(t - 1);
            }
            gemm2(T - 1);
            t_commit(bars + BAR_A2FULL);
        }
    } else {
        /* ===== 4 NQ warps, NQ threads per chain: candidate, link function, tail of the step =====
         * TMEM lane group lg = warp % 4 (hardware rule); the NQ warps of a lane group split everything in equal parts:
         * part h takes dimensions [DW h, DW (h + 1)) and tile columns [CW h, CW (h + 1)) */
        constexpr int DW = kTD / NQ, CW = kTRows / NQ;
        const int lg = warp & 3;
        const int h = (warp - 2) >> 2;
        const int m = 32 * lg + lane;                    /* chain of the CTA = TMEM lane = operand row */
        const long long ch = (long long)blockIdx.x * 128 + m;
        const bool active = ch < a.st.n;
        const long long pitch = a.st.pitch;
        const int cv = a.st.cv;
        const unsigned long long seed = active ? a.st.seeds[ch] : 0ull;
        const unsigned long long blk0 = a.step * amh::blocks_per_step_cv(cv, kTD);
        /* ---- candidate = x + (sigma z + drift grad)  (MALA.jl:70 -> proposal.jl:49-56), fp64, contract noise ---- */
        {
            double z[DW];                                /* the normals of dimensions DW h .. DW h + DW - 1 */
            if (cv == AMH_CONTRACT_V2) {
                for (int jb = 0; jb < DW / 4; ++jb) {
                    const amh::Block b = amh::stream_block7(seed, blk0 + (unsigned long long)(DW / 4 * h + jb), 0u);
                    amh::normal_quad(b, z[4 * jb], z[4 * jb + 1], z[4 * jb + 2], z[4 * jb + 3]);
                }
            } else {
                for (int jb = 0; jb < DW / 2; ++jb) {
                    const amh::Block b = amh::stream_block(seed, blk0 + (unsigned long long)(DW / 2 * h + jb), 0u);
                    amh::normal_pair(b, z[2 * jb], z[2 * jb + 1]);
                }
            }
            for (int j0 = 0; j0 < DW; j0 += 8) {
                unsigned hi[4], lo[4];
#pragma unroll
                for (int i = 0; i < 8; i += 2) {
                    double c[2] = {0.0, 0.0};
                    if (active) {
#pragma unroll
                        for (int e = 0; e < 2; ++e) {
                            const long long o = (long long)(DW * h + j0 + i + e) * pitch + ch;
                            c[e] = a.st.X[o] + (a.sigma * z[j0 + i + e] + a.drift * a.st.G[o]);
                            a.Xc[o] = c[e];
                        }
                    }
                    const __nv_bfloat162 hh = __floats2bfloat162_rn((float)c[0], (float)c[1]);
                    const float2 hf = __bfloat1622float2(hh);
                    const __nv_bfloat162 ll2 = __floats2bfloat162_rn((float)(c[0] - (double)hf.x), (float)(c[1] - (double)hf.y));
                    hi[i >> 1] = *reinterpret_cast<const unsigned*>(&hh);
                    lo[i >> 1] = *reinterpret_cast<const unsigned*>(&ll2);
                }
                const int dim0 = DW * h + j0;                                          /* K block = dim0 / 64, chunk = (dim0 % 64) / 8 */
                store_chunk(tsm + kOffC + (dim0 >> 6) * 16384, m, (dim0 & 63) >> 3, hi);
                store_chunk(tsm + kOffC + 32768 + (dim0 >> 6) * 16384, m, (dim0 & 63) >> 3, lo);
            }
        }
        t_fence_async_smem();
        __syncwarp();
        if (lane == 0) t_mbar_arrive(bars + BAR_CREADY);
        /* ---- per tile: eta -> (log-likelihood terms, residuals), columns CW h .. CW h + CW - 1 ---- */
        double ll = 0.0;
        /* the responses of a tile's rows come from L2: fetched one tile ahead (they were the top stall of the epilogue) */
        float ysn[CW];
        auto load_y = [&](int t) {
            int tq = t + trot; if (tq >= T) tq -= T;
            const float* yp = a.y + tq * kTRows + CW * h;
#pragma unroll
            for (int i = 0; i < CW; i += 4) {
                const float4 yv = __ldg(reinterpret_cast<const float4*>(yp + i));
                ysn[i] = yv.x; ysn[i + 1] = yv.y; ysn[i + 2] = yv.z; ysn[i + 3] = yv.w;
            }
        };
        load_y(0);
        for (int t = 0; t < T; ++t) {
            const int s = t & 1, u = t >> 1;
            float ysc[CW];
#pragma unroll
            for (int i = 0; i < CW; ++i) ysc[i] = ysn[i];
            if (t + 1 < T) load_y(t + 1);
            t_mbar_wait(bars + BAR_A1FULL + s, (unsigned)(u & 1));
            t_fence_after();
            /* the tile's CW columns of this thread, processed STAGE BY STAGE over 16 values at a time (straight-line code:
             * 16 independent dependency chains for the scheduler; the MUFU latencies overlap instead of adding up) */
            int tt = t + trot; if (tt >= T) tt -= T;
            const int row0 = tt * kTRows + CW * h;
            const int nlive = a.nrows - row0;                                   /* rows beyond the data are padding */
            unsigned rh[CW / 2], rl[CW / 2];
            float llt = 0.0f;
#pragma unroll
            for (int c0 = 0; c0 < CW; c0 += 16) {
                float eta[16];
                t_ld16(tmem + ((unsigned)(32 * lg) << 16) + (unsigned)(s * 64 + CW * h + c0), eta);
                float ys[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) ys[i] = ysc[c0 + i];
                float ex[16], w[16], iw[16], lw[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) ex[i] = ex2_approx(-1.4426950408889634f * fabsf(eta[i]));   /* e^-|eta| */
#pragma unroll
                for (int i = 0; i < 16; ++i) w[i] = 1.0f + ex[i];
#pragma unroll
                for (int i = 0; i < 16; ++i) { iw[i] = rcp_approx(w[i]); lw[i] = lg2_approx(w[i]); }
                /* rows beyond the data need no mask: their design-matrix rows and responses are zero, so eta = 0 exactly,
                 * their residual (-1/2) meets a zero row in GEMM2, and their log-likelihood term is -log1pexp(0) = -ln 2
                 * each, which is added back once per tile below */
                float rv[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const float l1p = fmaf(lw[i], 0.6931471805599453f, fmaxf(eta[i], 0.0f));    /* log1pexp(eta) */
                    const float sg = eta[i] >= 0.0f ? iw[i] : ex[i] * iw[i];                    /* sigmoid(eta)  */
                    llt += fmaf(ys[i], eta[i], -l1p);
                    rv[i] = ys[i] - sg;
                }
#pragma unroll
                for (int i = 0; i < 16; i += 2) {
                    const __nv_bfloat162 hh = __floats2bfloat162_rn(rv[i], rv[i + 1]);
                    const float2 hf = __bfloat1622float2(hh);
                    const __nv_bfloat162 l2 = __floats2bfloat162_rn(rv[i] - hf.x, rv[i + 1] - hf.y);
                    rh[(c0 + i) >> 1] = *reinterpret_cast<const unsigned*>(&hh);
                    rl[(c0 + i) >> 1] = *reinterpret_cast<const unsigned*>(&l2);
                }
            }
            /* only now does R have to be free: GEMM2 of the previous tile ran under the arithmetic above */
            t_mbar_wait(bars + BAR_REMPTY, (unsigned)((t & 1) ^ 1));
#pragma unroll
            for (int g = 0; g < CW / 8; ++g) {
                const unsigned eh[4] = {rh[4 * g], rh[4 * g + 1], rh[4 * g + 2], rh[4 * g + 3]};
                const unsigned el[4] = {rl[4 * g], rl[4 * g + 1], rl[4 * g + 2], rl[4 * g + 3]};
                store_chunk(tsm + kOffR, m, CW / 8 * h + g, eh);
                store_chunk(tsm + kOffR + 16384, m, CW / 8 * h + g, el);
            }
            const int npad = min(CW, max(0, CW - nlive));
            ll += (double)llt + (double)npad * 0.6931471805599453;
            t_fence_before();                            /* this thread's TMEM reads of the accumulator are done */
            t_fence_async_smem();                        /* its R rows are visible to the tensor core */
            __syncwarp();
            if (lane == 0) { t_mbar_arrive(bars + BAR_A1EMPTY + s); t_mbar_arrive(bars + BAR_RFULL); }
        }
        /* ---- tail of the step (fp64): gradient, Hastings terms, accept  (MALA.jl:73-93); dimensions DW h .. DW h + DW - 1 ---- */
        t_mbar_wait(bars + BAR_A2FULL, 0u);              /* every MMA has completed: R's shared memory is free too */
        t_fence_after();
        double q = 0.0, A = 0.0, Bq = 0.0;
#pragma unroll 1
        for (int cc = 0; cc < DW; cc += 16) {
            const int c0 = DW * h + cc;
            float gg[16];
            t_ld16(tmem + ((unsigned)(32 * lg) << 16) + (unsigned)(128 + c0), gg);
            if (active) {
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const long long o = (long long)(c0 + i) * pitch + ch;
                    const double c = a.Xc[o], x = a.st.X[o], g = a.st.G[o];
                    const double gc = (double)gg[i] - c * a.invtau2;
                    a.Gc[o] = gc;
                    q = fma(c, c, q);
                    const double da = (x - c) - a.drift * gc;
                    const double db = (c - x) - a.drift * g;
                    A = fma(da, da, A);
                    Bq = fma(db, db, Bq);
                }
            }
        }
        t_fence_before();
        /* the parts of a chain meet in shared memory (R's region is free now) */
        double* part = reinterpret_cast<double*>(tsm + kOffR);          /* [NQ - 1][128 chains][4]: ll, q, A, B of parts 1.. */
        int* verdict = reinterpret_cast<int*>(tsm + kOffR + (NQ - 1) * 4096);       /* [128 chains] */
        if (h >= 1) { double* pp = part + (size_t)(h - 1) * 512 + 4 * m; pp[0] = ll; pp[1] = q; pp[2] = A; pp[3] = Bq; }
        asm volatile("bar.sync 1, %0;" ::"n"(128 * NQ) : "memory");
        double lp = 0.0, lp_c = 0.0;
        bool acc = false;
        if (h == 0) {
            if (active) {
                lp = a.st.lp[ch];
                double llf = ll, qf = q, Af = A, Bf = Bq;
#pragma unroll
                for (int k = 0; k < NQ - 1; ++k) {
                    const double* pp = part + (size_t)k * 512 + 4 * m;
                    llf += pp[0]; qf += pp[1]; Af += pp[2]; Bf += pp[3];
                }
                lp_c = llf - qf * a.inv2tau2;
                const double logratio = (-0.5 * (Af / a.sigma2)) - (-0.5 * (Bf / a.sigma2));
                const double loga = (lp_c - lp) + logratio;
                const double e = amh::step_exponential_cv(cv, seed, blk0, kTD);
                acc = -e < loga;                                          /* MALA.jl:86 (strict) */
            }
            verdict[m] = acc ? 1 : 0;
        }
        asm volatile("bar.sync 1, %0;" ::"n"(128 * NQ) : "memory");
        acc = verdict[m] != 0;
        if (active) {
            if (acc) {
                for (int j = DW * h; j < DW * h + DW; ++j) {
                    const long long o = (long long)j * pitch + ch;
                    a.st.X[o] = a.Xc[o];
                    a.st.G[o] = a.Gc[o];
                }
            }
            if (a.sv.out || a.sv.sum) {
                for (int j = DW * h; j < DW * h + DW; ++j) {
                    const long long o = (long long)j * pitch + ch;
                    const double v = a.st.X[o];
                    if (a.sv.out) a.sv.out[(long long)j * a.sv.out_pitch + ch] = v;
                    if (a.sv.sum) save_moments(a.sv, o, v);
                }
            }
            if (h == 0) {
                if (acc) {
                    a.st.lp[ch] = lp_c;
                    a.st.nacc[ch] = a.st.nacc[ch] + 1ull;
                }
                a.st.acc[ch] = acc ? 1 : 0;
                if (a.sv.out) a.sv.out[(long long)kTD * a.sv.out_pitch + ch] = acc ? lp_c : lp;
                if (a.sv.acc_out) a.sv.acc_out[ch] = acc ? 1 : 0;
            }
        }
    }
    t_fence_before();
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256));
}

/* ---- operand preparation: X (fp64, row-major) -> bf16 slices, row-major and transposed, padded to whole tiles ---- */
__global__ void mala_tensor_prep_kernel(const double* X, const double* y, int nrows, int rows_pad, __nv_bfloat16* Xhi, __nv_bfloat16* Xlo,
                                        __nv_bfloat16* XThi, __nv_bfloat16* XTlo, float* y32) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < (long long)rows_pad * kTD) {
        const int row = (int)(idx / kTD), j = (int)(idx % kTD);
        const double v = row < nrows ? X[(long long)row * kTD + j] : 0.0;
        const __nv_bfloat16 h = __float2bfloat16_rn((float)v);
        const __nv_bfloat16 l = __float2bfloat16_rn((float)(v - (double)__bfloat162float(h)));
        Xhi[idx] = h; Xlo[idx] = l;
        XThi[(long long)j * rows_pad + row] = h; XTlo[(long long)j * rows_pad + row] = l;
    }
    if (idx < rows_pad) y32[idx] = idx < nrows ? (float)y[idx] : 0.0f;
}

namespace {
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_tiled() {
    static EncodeTiledFn fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess) p = nullptr;
        return (EncodeTiledFn)p;
    }();
    return fn;
}
/* 2-d bf16 tensor [outer][inner], box [box_outer][64] (64 elements = one 128-byte swizzle row) */
int make_map(CUtensorMap* m, const void* base, long long inner, long long outer, int box_outer) {
    EncodeTiledFn fn = encode_tiled();
    if (!fn) return fail(AMH_ERR_CUDA, "cuTensorMapEncodeTiled unavailable (cudaGetDriverEntryPoint)");
    cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)outer};
    cuuint64_t strides[1] = {(cuuint64_t)inner * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)box_outer};
    cuuint32_t estr[2] = {1, 1};
    const CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(AMH_ERR_CUDA, "cuTensorMapEncodeTiled failed");
    return AMH_OK;
}
}  // namespace

struct MalaTensorState {
    CUtensorMap maps[4];          /* X hi, X lo, X^T hi, X^T lo */
    int rows_pad = 0;
    float* y32 = nullptr;
    double* Xc = nullptr;
    double* Gc = nullptr;
};

bool mala_tensor_eligible(const amh_run& r) {
    return r.sampler->d.kind == AMH_SAMPLER_MALA && r.sampler->d.precision == AMH_PRECISION_BF16X2 && r.target->kind == AMH_TARGET_LOGISTIC &&
           r.dim == kTD && r.target->ndata >= 64;
}

void mala_tensor_release(amh_run& r) {
    delete (MalaTensorState*)r.tensor_state;
    r.tensor_state = nullptr;
}

int launch_mala_tensor(amh_run& r, int nsteps, const SaveArgs& sv) {
    const amh_sampler& s = *r.sampler;
    const amh_target& t = *r.target;
    const int nrows = (int)t.ndata;
    const size_t np = (size_t)r.pitch;
    if (!r.tensor_state) {
        MalaTensorState* st = new MalaTensorState();
        st->rows_pad = (nrows + kTRows - 1) / kTRows * kTRows;
        const size_t ne = (size_t)st->rows_pad * kTD;
        /* [Xhi | Xlo | XThi | XTlo] bf16, y32, Xc, Gc */
        const size_t bytes = 4 * ne * 2 + (size_t)st->rows_pad * 4 + 2 * (size_t)kTD * np * 8 + 1024;
        const int rca = dmalloc(r.ctx, &r.scratch2, bytes);
        if (rca) { delete st; return rca; }
        unsigned char* base = (unsigned char*)r.scratch2;
        __nv_bfloat16* Xhi = (__nv_bfloat16*)base;
        __nv_bfloat16* Xlo = Xhi + ne;
        __nv_bfloat16* XThi = Xlo + ne;
        __nv_bfloat16* XTlo = XThi + ne;
        st->y32 = (float*)(XTlo + ne);
        st->Xc = (double*)(((uintptr_t)(st->y32 + st->rows_pad) + 255) & ~(uintptr_t)255);
        st->Gc = st->Xc + (size_t)kTD * np;
        const unsigned grid = (unsigned)((ne + 255) / 256);
        mala_tensor_prep_kernel<<<grid, 256, 0, r.ctx->stream>>>(t.dblob + 1, t.dblob + 1 + (size_t)nrows * kTD, nrows, st->rows_pad, Xhi, Xlo, XThi,
                                                                  XTlo, st->y32);
        AMH_CUDA_TRY(cudaGetLastError());
        int rc = make_map(&st->maps[0], Xhi, kTD, st->rows_pad, kTRows);
        if (!rc) rc = make_map(&st->maps[1], Xlo, kTD, st->rows_pad, kTRows);
        if (!rc) rc = make_map(&st->maps[2], XThi, st->rows_pad, kTD, kTD);
        if (!rc) rc = make_map(&st->maps[3], XTlo, st->rows_pad, kTD, kTD);
        if (rc) { delete st; return rc; }
        AMH_CUDA_TRY(cudaFuncSetAttribute(mala_tensor_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemT));
        AMH_CUDA_TRY(cudaFuncSetAttribute(mala_tensor_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemT));
        r.tensor_state = st;
    }
    MalaTensorState& st = *(MalaTensorState*)r.tensor_state;
    MalaTArgs a;
    std::memset(&a, 0, sizeof(a));
    a.st = chain_state(r);
    a.sigma = s.mala_sigma; a.sigma2 = s.d.mala_sigma2; a.drift = s.d.mala_drift;
    a.inv2tau2 = t.inv2tau2; a.invtau2 = t.invtau2;
    a.nrows = nrows; a.ntiles = st.rows_pad / kTRows;
    a.y = st.y32; a.Xc = st.Xc; a.Gc = st.Gc;
    const unsigned grid = (unsigned)((r.n + 127) / 128);
    static const int nq = [] { const char* ev = std::getenv("AMH_K3T_NQ"); return (ev && std::atoi(ev) == 2) ? 2 : 4; }();   /* A/B switch */
    SaveArgs none;
    std::memset(&none, 0, sizeof(none));
    for (int k = 0; k < nsteps; ++k) {
        a.step = (unsigned long long)(r.step + k + 1);
        a.sv = (k == nsteps - 1) ? sv : none;
        if (nq == 4) mala_tensor_kernel<4><<<grid, 576, kSmemT, r.ctx->stream>>>(a, st.maps[0], st.maps[1], st.maps[2], st.maps[3]);
        else mala_tensor_kernel<2><<<grid, 320, kSmemT, r.ctx->stream>>>(a, st.maps[0], st.maps[1], st.maps[2], st.maps[3]);
        AMH_CUDA_TRY(cudaGetLastError());
        r.launches += 1;
        r.pending_launches += 1;
    }
    return AMH_OK;
}

}  // namespace amhh
