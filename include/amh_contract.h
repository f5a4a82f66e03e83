/* amh_contract.h -- the numerical CONTRACT shared by the sm_100a kernels and the
 * CPU oracle: counter-based random streams and deterministic fp64 math.
 *
 * Why this exists (SURVEY.md 7 "hard part 1", Appendix B): the reference draws
 * its noise from Julia's Xoshiro + ziggurat (`randn`/`randexp`/`rand`, called at
 * /root/reference/src/proposal.jl:25-28, mh-core.jl:108, emcee.jl:52,81,93,
 * MALA.jl:86, RobustAdaptiveMetropolis.jl:135,148,193), whose word consumption
 * is data dependent and not stable across Julia versions.  A lock-step SIMT
 * engine needs a fixed-slot counter stream, so parity is defined on THIS
 * contract: Philox4x32-10 keyed by the per-chain seed, a fixed word budget per
 * step, and transcendental functions built only from IEEE-754 correctly
 * rounded operations (+ - * / sqrt fma, integer/bit ops).  Host code must be
 * compiled with -ffp-contract=off and device code with -fmad=false; every
 * fused multiply-add below is an explicit fma().  Under those flags gcc and
 * nvcc produce bit-identical results, which is what lets tests demand
 * bit-exact states and accept/reject decisions between GPU and oracle.
 *
 * A future Julia `PhiloxRNG <: AbstractRNG` that is re-positioned to word
 * k*W at the start of step k reads the same values.
 */
#ifndef AMH_CONTRACT_H
#define AMH_CONTRACT_H

#include <stdint.h>
#include <string.h>
#include <math.h>
#include "amh_contract_tables.h"

#if defined(__CUDACC__)
#define AMH_HD __host__ __device__ __forceinline__
#else
#define AMH_HD static inline
#endif

/* Contract versions.  A run is created under ONE version (amh_sampler_desc.contract; 0 = AMH_CONTRACT_VERSION) and
 * keeps it for life; kernels and oracle implement both, bit for bit.
 *   v1  every normal PAIR consumes one Philox4x32-10 block: 64-bit uniforms for radius and angle.
 *   v2  the step noise of MH / MALA / RAM (the d normals and the exponential of a step, stream 0) comes from
 *       Philox4x32-7 blocks -- the Crush-resistant round count of Salmon et al. (SC'11, table 2) -- and one block
 *       yields FOUR normals (two Box-Muller pairs, 32-bit radius and angle words).  Why: on B200 the 32x32->64
 *       multiplies of Philox (IMAD.WIDE) issue on the FP64 datapath and cost a quarter of the d = 32 step
 *       (tools/ubench/issue_probe.cu); v2 needs 63 of them per 16 normals instead of 170.  Everything else
 *       -- the stretch move's partner / uniform / exponential words, the univariate families' sub-streams, initial
 *       ensemble draws, every transcendental function -- is v1's, unchanged. */
#define AMH_CONTRACT_V1 1
#define AMH_CONTRACT_V2 2
#define AMH_CONTRACT_VERSION 2   /* the default of new runs */

namespace amh {

/* ------------------------------------------------------------------ bits */
AMH_HD uint32_t hi32(double x) {
#if defined(__CUDA_ARCH__)
    return (uint32_t)__double2hiint(x);
#else
    uint64_t b; memcpy(&b, &x, 8); return (uint32_t)(b >> 32);
#endif
}
AMH_HD uint32_t lo32(double x) {
#if defined(__CUDA_ARCH__)
    return (uint32_t)__double2loint(x);
#else
    uint64_t b; memcpy(&b, &x, 8); return (uint32_t)b;
#endif
}
AMH_HD double make_double(uint32_t hi, uint32_t lo) {
#if defined(__CUDA_ARCH__)
    return __hiloint2double((int)hi, (int)lo);
#else
    uint64_t b = ((uint64_t)hi << 32) | lo; double x; memcpy(&x, &b, 8); return x;
#endif
}

/* ------------------------------------------------------------ Philox4x32-10
 * Salmon, Moraes, Dror, Shaw, "Parallel random numbers: as easy as 1, 2, 3"
 * (SC'11).  Pinned in tests against the Random123 known-answer vectors. */
struct Block { uint32_t v[4]; };

template <int ROUNDS>
AMH_HD Block philox4x32_r(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                          uint32_t k0, uint32_t k1) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
    const uint32_t W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int r = 0; r < ROUNDS; ++r) {
        const uint64_t p0 = (uint64_t)M0 * c0;
        const uint64_t p1 = (uint64_t)M1 * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        c1 = (uint32_t)p1; c3 = (uint32_t)p0; c0 = n0; c2 = n2;
        k0 += W0; k1 += W1;
    }
    Block b; b.v[0] = c0; b.v[1] = c1; b.v[2] = c2; b.v[3] = c3;
    return b;
}
AMH_HD Block philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
    return philox4x32_r<10>(c0, c1, c2, c3, k0, k1);
}
AMH_HD Block philox4x32_7(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
    return philox4x32_r<7>(c0, c1, c2, c3, k0, k1);
}

/* Block `blk` of the stream of the chain whose 64-bit seed is `seed`.
 * key = (lo32(seed), hi32(seed)); counter = (lo32(blk), hi32(blk), stream, 0). */
AMH_HD Block stream_block(uint64_t seed, uint64_t blk, uint32_t stream) {
    return philox4x32_10((uint32_t)blk, (uint32_t)(blk >> 32), stream, 0u,
                         (uint32_t)seed, (uint32_t)(seed >> 32));
}
/* the same with 7 rounds: the step-noise blocks of contract v2 */
AMH_HD Block stream_block7(uint64_t seed, uint64_t blk, uint32_t stream) {
    return philox4x32_7((uint32_t)blk, (uint32_t)(blk >> 32), stream, 0u,
                        (uint32_t)seed, (uint32_t)(seed >> 32));
}
/* step-noise block of a run under contract `cv` */
AMH_HD Block step_block(int cv, uint64_t seed, uint64_t blk) {
    return cv == AMH_CONTRACT_V2 ? stream_block7(seed, blk, 0u) : stream_block(seed, blk, 0u);
}

/* --------------------------------------------------------------- uniform
 * 64-bit word (lo,hi) -> u = ((w >> 12) + 1/2) * 2^-52, strictly in (0,1);
 * the value is an odd multiple of 2^-53, hence exact. */
AMH_HD double u01(uint32_t wlo, uint32_t whi) {
    const double d = make_double(0x3FF00000u | (whi >> 12), (whi << 20) | (wlo >> 12));
    return d - 0x1.fffffffffffffp-1;   /* d - (1 - 2^-53) */
}
/* 32-bit word -> u = (w + 1/2) * 2^-32, strictly in (0,1), exact (contract v2 radius) */
AMH_HD double u01_32(uint32_t w) {
    const double d = make_double(0x3FF00000u | (w >> 12), w << 20);
    return d - 0x1.ffffffffp-1;        /* d - (1 - 2^-33) */
}

/* ------------------------------------------------------------------- log
 * Table-driven: z = x / 2^k in [0.6875, 1.375), r = z/c_i - 1 with |r| < 2^-7,
 * ln x = k ln2 - ln(1/c_i) + ln(1+r).  The two intervals adjacent to 1 use
 * c = 1 exactly, so ln(u) keeps full relative accuracy for u -> 1.
 * Attribution: the scheme (offset 0x3FE6..., 128-entry invc / logc table, r = z * invc - 1, polynomial in r) is the one
 * of the ARM Optimized Routines / glibc double-precision log (Szabolcs Nagy, MIT / LGPL); a third-party algorithm, not
 * the reference's.  The table here is regenerated by tools/gen_contract_tables.py with mpmath, not copied. */
struct alignas(16) LogTabEntry { double invc, nlogc; };

#if defined(__CUDACC__)
static __device__ const LogTabEntry amh_log_tab_dev[128] = { AMH_LOG_TABLE_ENTRIES };
#endif
static const LogTabEntry amh_log_tab_host[128] = { AMH_LOG_TABLE_ENTRIES };

/* -ln(x) for a positive, finite, NORMAL double x (no special cases). */
AMH_HD double neglog_normal(double x) {
    const uint32_t hx = hi32(x);
    const uint32_t tmp = hx - 0x3FE60000u;
    const uint32_t i = (tmp >> 13) & 127u;
    const int32_t k = (int32_t)tmp >> 20;
    const double z = make_double(hx - (tmp & 0xFFF00000u), lo32(x));
#if defined(__CUDA_ARCH__)
    const double invc = amh_log_tab_dev[i].invc, nlogc = amh_log_tab_dev[i].nlogc;
#else
    const double invc = amh_log_tab_host[i].invc, nlogc = amh_log_tab_host[i].nlogc;
#endif
    const double r = fma(z, invc, -1.0);
    const double w = fma((double)k, AMH_NEG_LN2, nlogc);      /* -(k ln2 + ln c) */
    double p = fma(r, AMH_LOG_L5, AMH_LOG_L4);
    p = fma(r, p, AMH_LOG_L3);
    p = fma(r, p, AMH_LOG_L2);
    p = fma(r, p, AMH_LOG_L1);
    p = fma(r, p, AMH_LOG_L0);
    const double r2 = r * r;
    /* -ln x = w - r - r^2 p */
    return (w - r) - r2 * p;
}

/* ln(x) for any double, IEEE special cases included. */
AMH_HD double log_(double x) {
    const uint32_t hx = hi32(x);
    if (hx - 0x00100000u >= 0x7FE00000u) {          /* zero, subnormal, neg, inf, nan */
        if (x == 0.0) return -INFINITY;
        if (x != x) return x;
        if (hx & 0x80000000u) return NAN;
        if (hx >= 0x7FF00000u) return x;            /* +inf */
        /* subnormal: scale by 2^54 */
        return -(neglog_normal(x * 0x1p54) + 54.0 * AMH_LN2);
    }
    return -neglog_normal(x);
}

/* ------------------------------------------------------------------- exp
 * k = rint(x/ln2), r = x - k ln2 in two steps, e^r by a degree-13 Taylor
 * polynomial (|r| <= 0.3466 -> truncation < 2^-60), scaled by 2^k exactly. */
AMH_HD double exp_(double x) {
    if (x != x) return x;
    if (x > 709.782712893384) return INFINITY;
    if (x < -745.2) return 0.0;
    const double kd = rint(x * AMH_INV_LN2);
    const double r = fma(kd, -AMH_LN2_LO, fma(kd, -AMH_LN2_HI, x));
    double p = fma(r, AMH_EXP_E13, AMH_EXP_E12);
    p = fma(r, p, AMH_EXP_E11);
    p = fma(r, p, AMH_EXP_E10);
    p = fma(r, p, AMH_EXP_E9);
    p = fma(r, p, AMH_EXP_E8);
    p = fma(r, p, AMH_EXP_E7);
    p = fma(r, p, AMH_EXP_E6);
    p = fma(r, p, AMH_EXP_E5);
    p = fma(r, p, AMH_EXP_E4);
    p = fma(r, p, AMH_EXP_E3);
    p = fma(r, p, AMH_EXP_E2);
    const double r2 = r * r;
    const double y = fma(r2, p, r) + 1.0;            /* in [0.70, 1.42] */
    const int k = (int)kd;
    /* scale by 2^k in two exact steps so that subnormal results round once */
    const int k1 = k / 2, k2 = k - k1;
    const double s1 = make_double((uint32_t)(k1 + 1023) << 20, 0u);
    const double s2 = make_double((uint32_t)(k2 + 1023) << 20, 0u);
    return (y * s1) * s2;
}

/* log(1 + e^x), used by the logistic-regression target. */
AMH_HD double log1pexp(double x) {
    const double t = exp_(-fabs(x));                 /* in (0,1] */
    const double w = 1.0 + t;
    /* log1p(t) = log(w) + (t - (w - 1))/w  (first-order correction of the rounding of w) */
    const double l = log_(w) + (t - (w - 1.0)) / w;
    return (x > 0.0 ? x : 0.0) + l;
}

/* logistic sigmoid 1/(1+e^-x), written so that it never overflows */
AMH_HD double sigmoid(double x) {
    const double t = exp_(-fabs(x));
    const double s = 1.0 / (1.0 + t);
    return x >= 0.0 ? s : t * s;
}

/* ----------------------------------------------------------- exponential
 * e = -ln(u), u = u01(word).  Reference call sites: -randexp(rng) < loga
 * (mh-core.jl:108, MALA.jl:86), <= (emcee.jl:93), randexp > -loga (RAM :148). */
AMH_HD double exponential(uint32_t wlo, uint32_t whi) {
    return neglog_normal(u01(wlo, whi));
}

/* ------------------------------------------------------------ normal pair
 * Box-Muller on one Philox block: radius from word 0, angle from word 1.
 *   rad   = sqrt(-2 ln u01(w0))
 *   theta = (pi/2) (q + g),  q = top 2 bits of w1,  g = frac52(w1 >> 10) - 1/2
 *   (z0, z1) = rad (cos theta, sin theta)
 * sin/cos(pi/2 g) come from near-minimax polynomials on |g| <= 1/2; the
 * quadrant q is applied exactly by swap + sign flips. */
AMH_HD void box_muller(double u, uint32_t q, double g, double& z0, double& z1) {
    const double nl = neglog_normal(u);
    const double rad = sqrt(nl + nl);
    const double y = g * g;
    double s = fma(y, AMH_SIN_S6, AMH_SIN_S5);
    s = fma(y, s, AMH_SIN_S4);
    s = fma(y, s, AMH_SIN_S3);
    s = fma(y, s, AMH_SIN_S2);
    s = fma(y, s, AMH_SIN_S1);
    s = fma(y, s, AMH_SIN_S0);
    s = s * g;
    double c = fma(y, AMH_COS_C7, AMH_COS_C6);
    c = fma(y, c, AMH_COS_C5);
    c = fma(y, c, AMH_COS_C4);
    c = fma(y, c, AMH_COS_C3);
    c = fma(y, c, AMH_COS_C2);
    c = fma(y, c, AMH_COS_C1);
    c = fma(y, c, AMH_COS_C0);
    /* q=0:(c,s) 1:(-s,c) 2:(-c,-s) 3:(s,-c) */
    const bool swap = (q & 1u) != 0u;
    const double a = swap ? s : c;
    const double bb = swap ? c : s;
    const uint32_t sa = ((q + 1u) & 2u) << 30;       /* sign of first component  */
    const uint32_t sb = (q & 2u) << 30;              /* sign of second component */
    const double ca = make_double(hi32(a) ^ sa, lo32(a));
    const double cb = make_double(hi32(bb) ^ sb, lo32(bb));
    z0 = rad * ca;
    z1 = rad * cb;
}
AMH_HD void normal_pair(const Block& b, double& z0, double& z1) {
    const uint32_t whi = b.v[3], wlo = b.v[2];
    /* 52 fraction bits = bits 61..10 of the word */
    const double g = make_double(0x3FF00000u | ((whi >> 10) & 0x000FFFFFu),
                                 (whi << 22) | (wlo >> 10)) - 1.5;
    box_muller(u01(b.v[0], b.v[1]), whi >> 30, g, z0, z1);
}
/* contract v2: one Box-Muller pair from two 32-bit words (radius word wr, angle word wa: quadrant = top 2 bits,
 * g = its other 30 bits as a fraction, minus 1/2); the arithmetic after that is v1's */
AMH_HD void normal_pair32(uint32_t wr, uint32_t wa, double& z0, double& z1) {
    const double g = make_double(0x3FF00000u | ((wa >> 10) & 0x000FFFFFu), wa << 22) - 1.5;
    box_muller(u01_32(wr), wa >> 30, g, z0, z1);
}
/* contract v2: the four normals of one step-noise block: (v0, v1) -> z[0], z[1]; (v2, v3) -> z[2], z[3] */
AMH_HD void normal_quad(const Block& b, double& z0, double& z1, double& z2, double& z3) {
    normal_pair32(b.v[0], b.v[1], z0, z1);
    normal_pair32(b.v[2], b.v[3], z2, z3);
}

/* -------------------------------------------------------- bounded integer
 * floor(w * n / 2^64): uniform on 0..n-1 up to a bias of n/2^64.  Stands in
 * for rand(rng, Random.Sampler(rng, 1:n)) at emcee.jl:48,52. */
AMH_HD uint64_t bounded(uint32_t wlo, uint32_t whi, uint64_t n) {
    const uint64_t w = ((uint64_t)whi << 32) | wlo;
#if defined(__CUDA_ARCH__)
    return __umul64hi(w, n);
#else
    return (uint64_t)(((unsigned __int128)w * n) >> 64);
#endif
}

/* ------------------------------------------- univariate proposal families
 * Arrays of univariate distributions as proposals (proposal.jl:26-35: `map(rand, p.proposal)` and the
 * left-to-right sum of `logpdf(p_i, v_i)`; README.md:104-112 `StaticProposal([Normal(0,1), InverseGamma(2,3)])`,
 * test/emcee.jl:19 `StretchProposal([InverseGamma(2,3), Normal(0,1)])`).  Parametrisations are those of
 * Distributions.jl; `logc` is the normalising constant, computed once on the host and passed in, so that both
 * sides use the same bits.
 *
 * Stream use: component i of step k (k = 0: the initial draw) of a chain reads
 *   NORMAL / LOGNORMAL : the chain's regular Box-Muller slot z[i] of that step (stream 0)
 *   everything else    : blocks (blk = k*d + i, stream = 2, 3, ...) -- a private sub-stream, because the gamma
 *                        sampler is a rejection loop (attempt t: stream 2+2t -> normal, 3+2t -> uniform;
 *                        shape < 1 boost: stream 2 + 2*AMH_GAMMA_MAX_TRIES).
 * For the initial draw of an ensemble walker w the sub-stream block is w*d + i of the ENSEMBLE's key. */
#define AMH_FAM_NORMAL      1   /* Normal(p0 = mu, p1 = sigma)            logc = -log(sigma) - log(2pi)/2        */
#define AMH_FAM_INVGAMMA    2   /* InverseGamma(p0 = shape, p1 = scale)   logc = shape*log(scale) - lgamma(shape) */
#define AMH_FAM_GAMMA       3   /* Gamma(p0 = shape, p1 = scale)          logc = -shape*log(scale) - lgamma(shape) */
#define AMH_FAM_UNIFORM     4   /* Uniform(p0 = a, p1 = b)                logc = -log(b - a)                      */
#define AMH_FAM_EXPONENTIAL 5   /* Exponential(p0 = scale)                logc = -log(scale)                      */
#define AMH_FAM_LOGNORMAL   6   /* LogNormal(p0 = mu, p1 = sigma)         logc = -log(sigma) - log(2pi)/2        */
#define AMH_GAMMA_MAX_TRIES 32  /* acceptance >= 0.95 per try: the cap is reached with probability < 1e-41       */

AMH_HD bool family_uses_normal_slot(int fam) { return fam == AMH_FAM_NORMAL || fam == AMH_FAM_LOGNORMAL; }

/* Gamma(shape, 1) by Marsaglia & Tsang (2000), "A simple method for generating gamma variables" */
AMH_HD double gamma_mt(uint64_t seed, uint64_t blk, double shape) {
    const bool boost = shape < 1.0;
    const double a = boost ? shape + 1.0 : shape;
    const double dd = a - 0x1.5555555555555p-2;          /* a - 1/3 */
    const double c = 1.0 / sqrt(9.0 * dd);
    double g = dd;
    for (uint32_t t = 0; t < AMH_GAMMA_MAX_TRIES; ++t) {
        const Block bn = stream_block(seed, blk, 2u + 2u * t);
        double x, unused;
        normal_pair(bn, x, unused);
        const double v1 = c * x + 1.0;
        if (!(v1 > 0.0)) continue;
        const double v = (v1 * v1) * v1;
        const Block bu = stream_block(seed, blk, 3u + 2u * t);
        const double u = u01(bu.v[0], bu.v[1]);
        g = dd * v;
        const double rhs = ((0.5 * (x * x) + dd) - g) + dd * log_(v);
        if (-neglog_normal(u) < rhs) break;
    }
    if (boost) {
        const Block bb = stream_block(seed, blk, 2u + 2u * AMH_GAMMA_MAX_TRIES);
        const double u = u01(bb.v[0], bb.v[1]);
        g = g * exp_(-neglog_normal(u) / shape);
    }
    return g;
}

/* rand(rng, p_i): `z` is the chain's standard normal of slot i (ignored by the families with a sub-stream) */
AMH_HD double family_draw(int fam, double p0, double p1, double z, uint64_t seed, uint64_t blk) {
    switch (fam) {
    case AMH_FAM_NORMAL:    return p1 * z + p0;
    case AMH_FAM_LOGNORMAL: return exp_(p1 * z + p0);
    case AMH_FAM_UNIFORM: {
        const Block b = stream_block(seed, blk, 2u);
        return (p1 - p0) * u01(b.v[0], b.v[1]) + p0;
    }
    case AMH_FAM_EXPONENTIAL: {
        const Block b = stream_block(seed, blk, 2u);
        return p0 * exponential(b.v[0], b.v[1]);
    }
    case AMH_FAM_GAMMA:     return p1 * gamma_mt(seed, blk, p0);
    case AMH_FAM_INVGAMMA:  return p1 / gamma_mt(seed, blk, p0);
    }
    return NAN;
}

/* logpdf(p_i, x), -Inf outside the support (what makes a random walk with a positive-only increment law reject) */
AMH_HD double family_logpdf(int fam, double p0, double p1, double logc, double x) {
    switch (fam) {
    case AMH_FAM_NORMAL: {
        const double z = (x - p0) / p1;
        return logc - 0.5 * (z * z);
    }
    case AMH_FAM_LOGNORMAL: {
        if (!(x > 0.0)) return x != x ? x : -INFINITY;
        const double lx = log_(x);
        const double z = (lx - p0) / p1;
        return (logc - lx) - 0.5 * (z * z);
    }
    case AMH_FAM_UNIFORM:     return (x >= p0 && x <= p1) ? logc : (x != x ? x : -INFINITY);
    case AMH_FAM_EXPONENTIAL: return (x >= 0.0) ? logc - x / p0 : (x != x ? x : -INFINITY);
    case AMH_FAM_GAMMA:
        if (!(x > 0.0)) return x != x ? x : -INFINITY;
        return (logc + (p0 - 1.0) * log_(x)) - x / p1;
    case AMH_FAM_INVGAMMA:
        if (!(x > 0.0)) return x != x ? x : -INFINITY;
        return (logc - (p0 + 1.0) * log_(x)) - p1 / x;
    }
    return NAN;
}

/* ----------------------------------------------------- stream word budget
 * Step k of a chain owns blocks [k*B, (k+1)*B) of its stream (k = 0 is the
 * initial draw).  For the d-dimensional MH / MALA / RAM steps
 *   B = ceil(d/2) + 1 : blocks 0..ceil(d/2)-1 -> normals z[2j], z[2j+1]
 *                       block  ceil(d/2)      -> word 0: the exponential
 * For the stretch move every walker move owns 2 blocks of its ENSEMBLE's
 * stream: block 0 word 0 -> partner index, word 1 -> uniform for z;
 *         block 1 word 0 -> exponential. */
AMH_HD uint64_t blocks_per_step(int d) { return (uint64_t)((d + 1) / 2 + 1); }

/* the same under contract `cv`: v2 packs four normals into a block, B = ceil(d/4) + 1, block ceil(d/4) word 0 (64 bits)
 * -> the exponential */
AMH_HD int normal_blocks(int cv, int d) { return cv == AMH_CONTRACT_V2 ? (d + 3) / 4 : (d + 1) / 2; }
AMH_HD uint64_t blocks_per_step_cv(int cv, int d) { return (uint64_t)(normal_blocks(cv, d) + 1); }

/* z[0..d-1] of the step whose first block is blk0 (generic form: one block after the other) */
AMH_HD void step_normals_cv(int cv, uint64_t seed, uint64_t blk0, int d, double* z) {
    if (cv == AMH_CONTRACT_V2) {
        for (int j = 0; 4 * j < d; ++j) {
            const Block b = stream_block7(seed, blk0 + (uint64_t)j, 0u);
            double q0, q1, q2, q3;
            normal_quad(b, q0, q1, q2, q3);
            z[4 * j] = q0;
            if (4 * j + 1 < d) z[4 * j + 1] = q1;
            if (4 * j + 2 < d) z[4 * j + 2] = q2;
            if (4 * j + 3 < d) z[4 * j + 3] = q3;
        }
    } else {
        for (int j = 0; 2 * j < d; ++j) {
            const Block b = stream_block(seed, blk0 + (uint64_t)j, 0u);
            double z0, z1;
            normal_pair(b, z0, z1);
            z[2 * j] = z0;
            if (2 * j + 1 < d) z[2 * j + 1] = z1;
        }
    }
}
/* the exponential of that step */
AMH_HD double step_exponential_cv(int cv, uint64_t seed, uint64_t blk0, int d) {
    const Block b = step_block(cv, seed, blk0 + (uint64_t)normal_blocks(cv, d));
    return exponential(b.v[0], b.v[1]);
}

}  /* namespace amh */

#endif /* AMH_CONTRACT_H */
