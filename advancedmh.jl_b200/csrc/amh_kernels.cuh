/* amh_kernels.cuh -- the fused per-step kernels (sm_100a).
 *
 * One thread owns one chain; chains are the fastest-varying index of every
 * state array ([dim][nchains]), so each warp-wide access is one contiguous
 * 256-byte request.  A launch fuses `nsteps` consecutive MCMC steps of every
 * chain; nsteps == 1 is exactly "one kernel per MCMC step".
 *
 *   K1  mh_step_kernel      StaticMH / RWMH          mh-core.jl:92-117
 *   K6  init_kernel         first step of any sampler mh-core.jl:76-86, emcee.jl:29-34,
 *                                                     MALA.jl:38-40, RAM :175-214
 */
#pragma once
#include "amh_device.cuh"

namespace amhd {

template <int DMAX>
struct MhArgs {
    ChainState st;
    SaveArgs sv;
    int d;
    int is_rw;                   /* RandomWalkProposal (1) or StaticProposal (0) */
    int hast;                    /* 0: Hastings term is exactly 0; 1: static, cached logq; 2: RW with non-zero mean */
    int nsteps;
    int pace;                    /* optional pause per step in ns (0 = never taken), see the step loop */
    unsigned long long step0;    /* stateful steps already taken */
    PropP<DMAX> prop;
};

/* ---------------------------------------------------------------------------
 * K1: fused Metropolis-Hastings step(s).
 * Per step and chain: Philox blocks -> Box-Muller normals -> proposal transform
 * (L z from the constant bank on the fixed-DMAX path) -> target log-density ->
 * Hastings correction -> exponential draw -> accept/reject.  The current state
 * lives in shared memory ([i][thread], conflict free), the candidate in
 * registers. */
/* HAST1 (exact dimensions, amh_launch_mh_hast.cu): StaticProposal with issymmetric = false -- the Hastings term
 * logq(state) - logq(candidate) with logq(state) cached per chain */
template <int DMAX, class T, int BLOCK, int MINB, bool HAST1 = false>
__global__ void __launch_bounds__(BLOCK, MINB)
mh_step_kernel(const __grid_constant__ MhArgs<DMAX> a,
               const __grid_constant__ typename T::template Params<DMAX> tp) {
    using D = Dim<DMAX>;
    constexpr int CAP = D::cap;
    constexpr int UNR = D::unr;
    extern __shared__ double sx[];
    const int tid = threadIdx.x;
    const long long ch = (long long)blockIdx.x * BLOCK + tid;
    if (ch >= a.st.n) return;
    const int d = D::fixed ? DMAX : a.d;      /* fixed path: launched only with dim == DMAX, every `i < d` folds */
    const int top = D::fixed ? DMAX : d;
    const unsigned long long seed = a.st.seeds[ch];
    double lp = a.st.lp[ch];
    double lq = (a.hast == 1) ? a.st.lq[ch] : 0.0;
    unsigned long long nacc = a.st.nacc[ch];
    unsigned char accepted = a.st.acc[ch];
#pragma unroll UNR
    for (int i = 0; i < top; ++i)
        if (i < d) sx[i * BLOCK + tid] = a.st.X[(long long)i * a.st.pitch + ch];

    const int cv = a.st.cv;
    const unsigned long long B = amh::blocks_per_step_cv(cv, d);
    double z[CAP];
    for (int s = 0; s < a.nsteps; ++s) {
        /* untaken by default, kept on purpose: with this branch at the top of the step ptxas orders the loop body
         * differently and the kernel is 4-5 % faster (d = 10: 1.76e10 -> 1.85e10, d = 2: 6.86e10 -> 7.15e10 chain-steps/s,
         * A/B of two builds on one box; the same effect as in K1T16) */
        if (a.pace > 0) __nanosleep((unsigned)a.pace);
        const unsigned long long k = a.step0 + (unsigned long long)s + 1ull;
        const unsigned long long blk0 = k * B;
        double e;
        if constexpr (D::fixed) {
            step_noise_fixed_cv<DMAX>(cv, seed, k, z, e);
        } else {
            step_normals<DMAX>(cv, seed, blk0, d, z);
            e = amh::step_exponential_cv(cv, seed, blk0, d);
        }
        draw_inplace<DMAX>(z, d, a.prop);
        if (a.is_rw) {
#pragma unroll UNR
            for (int i = 0; i < top; ++i)
                if (i < d) z[i] = sx[i * BLOCK + tid] + z[i];
        }
        const double lp_c = T::template logp<DMAX>(z, d, tp);
        double logratio = 0.0, lq_c = 0.0;
        if constexpr (!D::fixed) {
            if (a.hast == 1) {
                lq_c = logq<DMAX>(z, d, a.prop);
                logratio = lq - lq_c;
            } else if (a.hast == 2) {
                double t1[CAP], t2[CAP];
                for (int i = 0; i < d; ++i) {
                    const double xi = sx[i * BLOCK + tid];
                    t1[i] = xi - z[i];
                    t2[i] = z[i] - xi;
                }
                logratio = logq<DMAX>(t1, d, a.prop) - logq<DMAX>(t2, d, a.prop);
            }
        } else if constexpr (HAST1) {
            lq_c = logq_fixed<DMAX>(z, a.prop);
            logratio = lq - lq_c;
        }
        const double loga = (lp_c - lp) + logratio;
        if (-e < loga) {
#pragma unroll UNR
            for (int i = 0; i < top; ++i)
                if (i < d) sx[i * BLOCK + tid] = z[i];
            lp = lp_c;
            lq = lq_c;
            accepted = 1;
            ++nacc;
        } else {
            accepted = 0;
        }
    }

#pragma unroll UNR
    for (int i = 0; i < top; ++i) {
        if (i < d) {
            const double v = sx[i * BLOCK + tid];
            const long long o = (long long)i * a.st.pitch + ch;
            a.st.X[o] = v;
            if (a.sv.out) a.sv.out[(long long)i * a.sv.out_pitch + ch] = v;
            if (a.sv.sum) {
                save_moments(a.sv, o, v);
            }
        }
    }
    a.st.lp[ch] = lp;
    if (a.hast == 1) a.st.lq[ch] = lq;
    a.st.nacc[ch] = nacc;
    a.st.acc[ch] = accepted;
    if (a.sv.out) a.sv.out[(long long)d * a.sv.out_pitch + ch] = lp;
    if (a.sv.acc_out) a.sv.acc_out[ch] = accepted;
}

/* ---------------------------------------------------------------------------
 * K6: first step.  Always the generic (runtime-dim) instantiation: it runs once
 * per run and must give the same bits as the fixed path, which it does because
 * both are the same source with the same operation order. */
struct InitArgs {
    ChainState st;
    int d;
    int mode;          /* 0: X given; 1: draw from proposal, chain stream step 0;
                          2: randn, chain stream step 0 (RAM);
                          3: draw from proposal, ENSEMBLE stream 1 (stretch) */
    int want_grad;     /* MALA */
    int want_lq;       /* static MH, non-symmetric */
    int init_acc;      /* initial Transition.accepted (RAM: true) */
    long long n_walkers;
    PropP<0> prop;
    /* RAM */
    double* S;         /* [tri][pitch] or NULL */
    const double* S0;  /* packed lower or NULL = identity */
    const amh_component* comps;   /* array of univariate laws (proposal.jl:26-28, 132-140) or NULL */
};

template <class T>
__global__ void __launch_bounds__(64)
init_kernel(const __grid_constant__ InitArgs a, const __grid_constant__ typename T::template Params<0> tp) {
    constexpr int CAP = Dim<0>::cap;
    const long long ch = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (ch >= a.st.n) return;
    const int d = a.d;
    double x[CAP];
    if (a.mode == 0) {
        for (int i = 0; i < d; ++i) x[i] = a.st.X[(long long)i * a.st.pitch + ch];
    } else if (a.mode == 3) {
        const long long en = ch / a.n_walkers, w = ch % a.n_walkers;
        const unsigned long long seed = a.st.seeds[en];
        const unsigned long long B1 = (unsigned long long)((d + 1) / 2);
        for (int j = 0; 2 * j < d; ++j) {
            const amh::Block b = amh::stream_block(seed, (unsigned long long)w * B1 + j, 1u);
            double z0, z1;
            amh::normal_pair(b, z0, z1);
            x[2 * j] = z0;
            if (2 * j + 1 < CAP) x[2 * j + 1] = z1;
        }
        if (a.comps) draw_components(x, d, a.comps, seed, (unsigned long long)w * (unsigned long long)d);
        else draw_inplace<0>(x, d, a.prop);
    } else {
        step_normals<0>(a.st.cv, a.st.seeds[ch], 0ull, d, x);
        if (a.mode == 1) {
            if (a.comps) draw_components(x, d, a.comps, a.st.seeds[ch], 0ull);
            else draw_inplace<0>(x, d, a.prop);
        }
    }
    if (a.mode != 0)
        for (int i = 0; i < d; ++i) a.st.X[(long long)i * a.st.pitch + ch] = x[i];
    if (a.want_grad) {
        double lp, g[CAP];
        T::template logp_grad<0>(x, d, tp, lp, g);
        a.st.lp[ch] = lp;
        for (int i = 0; i < d; ++i) a.st.G[(long long)i * a.st.pitch + ch] = g[i];
    } else {
        a.st.lp[ch] = T::template logp<0>(x, d, tp);
    }
    if (a.want_lq) a.st.lq[ch] = logq<0>(x, d, a.prop);
    if (a.S) {
        for (int i = 0; i < d; ++i)
            for (int j = 0; j <= i; ++j)
                a.S[(long long)tri(i, j) * a.st.pitch + ch] = a.S0 ? __ldg(a.S0 + tri(i, j)) : (i == j ? 1.0 : 0.0);
    }
    if (a.init_acc >= 0) a.st.acc[ch] = (unsigned char)a.init_acc;
}

}  /* namespace amhd */
