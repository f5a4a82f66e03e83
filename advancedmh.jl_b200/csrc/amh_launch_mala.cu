/* amh_launch_mala.cu -- K3: fused MALA step (MALA.jl:54-93) for the proposal
 * closure g -> MvNormal(drift * g, sigma2 * I)  (README.md:180, test/runtests.jl:292,351).
 *
 * One thread per chain.  State (x, lp, grad) is cached exactly like the
 * reference's GradientTransition (MALA.jl:14-19): one value-and-gradient
 * evaluation per step.  Current x and grad live in shared memory
 * ([i][thread]); the candidate and its gradient in registers / local memory.
 *
 * The many-row logistic-regression configuration (d = 128, 10^4 rows) runs
 * through this kernel's generic instantiation with the scalar per-chain target
 * (TLogistic); its dense tiled variant is the next step of that row. */
#include <cstdlib>
#include "amh_params.cuh"

namespace amhh {
using namespace amhd;

/* @rtc-begin: the device code from here to @rtc-end is also compiled by NVRTC for user-supplied targets (amh_rtc.cu) */
struct MalaArgs {
    ChainState st;
    SaveArgs sv;
    int d;
    int nsteps;
    unsigned long long step0;
    double sigma, sigma2, drift;
};

template <int DMAX, class T, int BLOCK>
__global__ void __launch_bounds__(BLOCK)
mala_step_kernel(const __grid_constant__ MalaArgs a, const __grid_constant__ typename T::template Params<DMAX> tp) {
    using D = Dim<DMAX>;
    constexpr int CAP = D::cap;
    constexpr int UNR = D::unr;
    extern __shared__ double smem[];
    const int tid = threadIdx.x;
    const long long ch = (long long)blockIdx.x * BLOCK + tid;
    if (ch >= a.st.n) return;
    const int d = D::fixed ? DMAX : a.d;
    const int top = D::fixed ? DMAX : d;
    double* sx = smem;
    double* sg = smem + (size_t)d * BLOCK;
    const unsigned long long seed = a.st.seeds[ch];
    double lp = a.st.lp[ch];
    unsigned long long nacc = a.st.nacc[ch];
    unsigned char accepted = a.st.acc[ch];
#pragma unroll UNR
    for (int i = 0; i < top; ++i)
        if (i < d) {
            sx[i * BLOCK + tid] = a.st.X[(long long)i * a.st.pitch + ch];
            sg[i * BLOCK + tid] = a.st.G[(long long)i * a.st.pitch + ch];
        }
    const int cv = a.st.cv;
    const unsigned long long B = amh::blocks_per_step_cv(cv, d);
    double c[CAP], gc[CAP];
    for (int s = 0; s < a.nsteps; ++s) {
        const unsigned long long k = a.step0 + (unsigned long long)s + 1ull;
        const unsigned long long blk0 = k * B;
        double e;
        if constexpr (D::fixed) {
            step_noise_fixed_cv<DMAX>(cv, seed, k, c, e);
        } else {
            step_normals<DMAX>(cv, seed, blk0, d, c);
            e = amh::step_exponential_cv(cv, seed, blk0, d);
        }
        /* candidate = state + rand(MvNormal(drift*grad, sigma2*I))  (MALA.jl:70 -> proposal.jl:49-56) */
#pragma unroll UNR
        for (int i = 0; i < top; ++i)
            if (i < d) c[i] = sx[i * BLOCK + tid] + (a.sigma * c[i] + a.drift * sg[i * BLOCK + tid]);
        double lp_c;
        T::template logp_grad<DMAX>(c, d, tp, lp_c, gc);
        /* q(prop(grad_c), state, cand) - q(prop(grad), cand, state)  (MALA.jl:78-80) */
        double A = 0.0, Bq = 0.0;
#pragma unroll UNR
        for (int i = 0; i < top; ++i)
            if (i < d) {
                const double xi = sx[i * BLOCK + tid];
                const double da = (xi - c[i]) - a.drift * gc[i];
                const double db = (c[i] - xi) - a.drift * sg[i * BLOCK + tid];
                A = (i == 0) ? da * da : fma(da, da, A);
                Bq = (i == 0) ? db * db : fma(db, db, Bq);
            }
        const double logratio = (-0.5 * (A / a.sigma2)) - (-0.5 * (Bq / a.sigma2));
        const double loga = (lp_c - lp) + logratio;
        if (-e < loga) {                                   /* MALA.jl:86 */
#pragma unroll UNR
            for (int i = 0; i < top; ++i)
                if (i < d) {
                    sx[i * BLOCK + tid] = c[i];
                    sg[i * BLOCK + tid] = gc[i];
                }
            lp = lp_c;
            accepted = 1;
            ++nacc;
        } else {
            accepted = 0;
        }
    }
#pragma unroll UNR
    for (int i = 0; i < top; ++i)
        if (i < d) {
            const double v = sx[i * BLOCK + tid];
            const long long o = (long long)i * a.st.pitch + ch;
            a.st.X[o] = v;
            a.st.G[o] = sg[i * BLOCK + tid];
            if (a.sv.out) a.sv.out[(long long)i * a.sv.out_pitch + ch] = v;
            if (a.sv.sum) {
                save_moments(a.sv, o, v);
            }
        }
    a.st.lp[ch] = lp;
    a.st.nacc[ch] = nacc;
    a.st.acc[ch] = accepted;
    if (a.sv.out) a.sv.out[(long long)d * a.sv.out_pitch + ch] = lp;
    if (a.sv.acc_out) a.sv.acc_out[ch] = accepted;
}

/* @rtc-end */
template <int DMAX, class T>
int launch_mala_t(amh_run& r, int nsteps, const SaveArgs& sv) {
    constexpr int BLOCK = (DMAX == 0) ? 32 : 128;
    const amh_sampler& s = *r.sampler;
    MalaArgs a;
    std::memset(&a, 0, sizeof(a));
    a.st = chain_state(r);
    a.sv = sv;
    a.d = r.dim;
    a.nsteps = nsteps;
    a.step0 = (unsigned long long)r.step;
    a.sigma = s.mala_sigma; a.sigma2 = s.d.mala_sigma2; a.drift = s.d.mala_drift;
    const auto tp = make_tp<T, DMAX>(*r.target);
    const size_t smem = 2 * (size_t)r.dim * BLOCK * sizeof(double);
    const unsigned grid = (unsigned)((r.n + BLOCK - 1) / BLOCK);
    if constexpr (T::kind == AMH_TARGET_USER) {
        static_assert(DMAX == 0 && BLOCK == 32, "RK_MALA names mala_step_kernel<0, TUser, 32>");
        void* params[] = {(void*)&a, (void*)&tp};
        const int rc = rtc_launch(r, RK_MALA, grid, BLOCK, smem, params);
        if (rc) return rc;
    } else {
        auto kern = mala_step_kernel<DMAX, T, BLOCK>;
        if (smem > 48 * 1024) AMH_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<grid, BLOCK, smem, r.ctx->stream>>>(a, tp);
        AMH_CUDA_TRY(cudaGetLastError());
    }
    r.launches += 1;
    r.pending_launches += 1;
    return AMH_OK;
}

#ifndef AMH_MALA_EXTRA_TU
template <class T>
int launch_mala_dim(amh_run& r, int nsteps, const SaveArgs& sv) {
    {   /* the second translation unit's exact dimensions (amh_launch_mala_dims.cu) */
        bool taken = false;
        const int rc = launch_mala_more_dims(r, nsteps, sv, taken);
        if (taken) return rc;
    }
    switch (r.dim) {          /* exact-dimension instantiations; everything else is generic */
    case 2: return launch_mala_t<2, T>(r, nsteps, sv);
    case 3: return launch_mala_t<3, T>(r, nsteps, sv);
    case 4: return launch_mala_t<4, T>(r, nsteps, sv);
    case 5: return launch_mala_t<5, T>(r, nsteps, sv);
    case 8: return launch_mala_t<8, T>(r, nsteps, sv);
    case 10: return launch_mala_t<10, T>(r, nsteps, sv);
    case 16: return launch_mala_t<16, T>(r, nsteps, sv);
    }
    return launch_mala_t<0, T>(r, nsteps, sv);
}

int launch_mala(amh_run& r, int nsteps, const SaveArgs& sv) {
    static const bool scalar_only = std::getenv("AMH_MALA_PATH") && std::strcmp(std::getenv("AMH_MALA_PATH"), "scalar") == 0;
    if (nsteps > 0 && mala_tensor_eligible(r)) return launch_mala_tensor(r, nsteps, sv);      /* opt-in: precision bf16x2 */
    if (!scalar_only && mala_logistic_eligible(r)) return launch_mala_logistic(r, nsteps, sv);
    switch (r.target->kind) {
    case AMH_TARGET_MVNORMAL: return launch_mala_dim<TMvNormal>(r, nsteps, sv);
    case AMH_TARGET_GAUSS_PREC: return launch_mala_dim<TGaussPrec>(r, nsteps, sv);
    case AMH_TARGET_ROSENBROCK: return launch_mala_dim<TRosenbrock>(r, nsteps, sv);
    case AMH_TARGET_IID_NORMAL: return launch_mala_t<2, TIidNormal>(r, nsteps, sv);
    case AMH_TARGET_LOGISTIC: return launch_mala_t<0, TLogistic>(r, nsteps, sv);
    case AMH_TARGET_USER:
        if (r.target->has_grad()) return launch_mala_t<0, TUser>(r, nsteps, sv);
        break;
    }
    return fail(AMH_ERR_INVALID, "The gradient of the log density function is not defined");
}
#endif  /* AMH_MALA_EXTRA_TU */

}  // namespace amhh
