/* amh_launch_mh.cu -- instantiation + dispatch of K1 (StaticMH / RWMH step) and
 * K6 (first step).  Dispatch is by (target kind, dimension bucket); the
 * Hastings-term variants that need a second vector (hast != 0) and dimensions
 * above 32 go to the generic instantiation (DMAX == 0). */
#include "amh_params.cuh"

namespace amhh {
using namespace amhd;

template <int DMAX>
struct LaunchCfg {
    /* DMAX == 32: 64-thread CTAs, 7 per SM -> 65 536 chains are ONE wave on 148 SMs
     * (1024 CTAs <= 1036 slots), register budget 65536/(7*64) = 146 */
    static constexpr int block = (DMAX == 0 || DMAX >= 32) ? 64 : 128;
    static constexpr int minb = (DMAX == 0) ? 4 : (DMAX >= 32 ? 7 : (DMAX >= 16 ? 6 : 8));
};

template <int DMAX, class T>
int launch_mh_t(amh_run& r, int nsteps, const SaveArgs& sv) {
    constexpr int BLOCK = LaunchCfg<DMAX>::block;
    constexpr int MINB = LaunchCfg<DMAX>::minb;
    const amh_sampler& s = *r.sampler;
    MhArgs<DMAX> a;
    std::memset(&a, 0, sizeof(a));
    a.st = chain_state(r);
    a.sv = sv;
    a.d = r.dim;
    a.is_rw = s.d.kind == AMH_SAMPLER_RW;
    a.hast = 0;
    if (!s.d.symmetric) {
        if (s.d.kind == AMH_SAMPLER_STATIC) a.hast = 1;
        else if (s.has_mean) a.hast = 2;
    }
    a.nsteps = nsteps;
    a.step0 = (unsigned long long)r.step;
    a.prop = make_prop<DMAX>(s);
    const auto tp = make_tp<T, DMAX>(*r.target);
    const size_t smem = (size_t)r.dim * BLOCK * sizeof(double);
    const unsigned grid = (unsigned)((r.n + BLOCK - 1) / BLOCK);
    if constexpr (T::kind == AMH_TARGET_USER) {
        static_assert(DMAX == 0 && BLOCK == 64 && MINB == 4, "RK_MH names mh_step_kernel<0, TUser, 64, 4>");
        void* params[] = {(void*)&a, (void*)&tp};
        const int rc = rtc_launch(r, RK_MH, grid, BLOCK, smem, params);
        if (rc) return rc;
    } else {
        auto kern = mh_step_kernel<DMAX, T, BLOCK, MINB>;
        if (smem > 48 * 1024) AMH_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<grid, BLOCK, smem, r.ctx->stream>>>(a, tp);
        AMH_CUDA_TRY(cudaGetLastError());
    }
    r.launches += 1;
    r.pending_launches += 1;
    return AMH_OK;
}

static bool needs_generic(const amh_run& r) {
    const amh_sampler& s = *r.sampler;
    if (r.dim > 32) return true;
    if (!s.d.symmetric && (s.d.kind == AMH_SAMPLER_STATIC || s.has_mean)) return true;
    return false;
}
/* StaticProposal with issymmetric = false (the reference's default for StaticMH): exact-dimension kernels with the
 * Hastings term live in amh_launch_mh_hast.cu */
static bool wants_hast1(const amh_run& r) {
    const amh_sampler& s = *r.sampler;
    return r.dim <= 32 && !s.d.symmetric && s.d.kind == AMH_SAMPLER_STATIC && !s.by_components();
}

#ifndef AMH_MH_EXTRA_TU
template <class T>
int launch_mh_full(amh_run& r, int nsteps, const SaveArgs& sv) {
    if (needs_generic(r)) return launch_mh_t<0, T>(r, nsteps, sv);
    {   /* the second translation unit's exact dimensions (amh_launch_mh_dims.cu) */
        bool taken = false;
        const int rc = launch_mh_more_dims(r, nsteps, sv, taken);
        if (taken) return rc;
    }
    switch (r.dim) {          /* exact-dimension instantiations; everything else is generic */
    case 1: return launch_mh_t<1, T>(r, nsteps, sv);
    case 2: return launch_mh_t<2, T>(r, nsteps, sv);
    case 3: return launch_mh_t<3, T>(r, nsteps, sv);
    case 4: return launch_mh_t<4, T>(r, nsteps, sv);
    case 5: return launch_mh_t<5, T>(r, nsteps, sv);
    case 6: return launch_mh_t<6, T>(r, nsteps, sv);
    case 8: return launch_mh_t<8, T>(r, nsteps, sv);
    case 10: return launch_mh_t<10, T>(r, nsteps, sv);
    case 12: return launch_mh_t<12, T>(r, nsteps, sv);
    case 16: return launch_mh_t<16, T>(r, nsteps, sv);
    case 20: return launch_mh_t<20, T>(r, nsteps, sv);
    case 24: return launch_mh_t<24, T>(r, nsteps, sv);
    case 32: return launch_mh_t<32, T>(r, nsteps, sv);
    }
    return launch_mh_t<0, T>(r, nsteps, sv);
}

template <class T>
int launch_mh_d2(amh_run& r, int nsteps, const SaveArgs& sv) {
    if (needs_generic(r)) return launch_mh_t<0, T>(r, nsteps, sv);
    return launch_mh_t<2, T>(r, nsteps, sv);
}

int launch_mh(amh_run& r, int nsteps, const SaveArgs& sv) {
    if (wants_hast1(r)) {
        bool taken = false;
        const int rc = launch_mh_hast(r, nsteps, sv, taken);
        if (taken) return rc;
    }
    if (r.mh_path != 1 && mh_logistic_eligible(r)) return launch_mh_logistic(r, nsteps, sv);
    if (r.mh_path != 1 && mh_tc_small_eligible(r)) return launch_mh_tc_padded(r, nsteps, sv);
    if (r.mh_path != 1 && mh_tc_eligible(r)) return launch_mh_tc(r, nsteps, sv);
    if (r.mh_path != 1 && mh_tc_padded_eligible(r)) return launch_mh_tc_padded(r, nsteps, sv);
    if (r.dim > kGenericCap) return fail(AMH_ERR_UNSUPPORTED, "StaticMH/RWMH on the device supports dim <= 128");
    switch (r.target->kind) {
    case AMH_TARGET_MVNORMAL: return launch_mh_full<TMvNormal>(r, nsteps, sv);
    case AMH_TARGET_GAUSS_PREC: return launch_mh_full<TGaussPrec>(r, nsteps, sv);
    case AMH_TARGET_ROSENBROCK: return launch_mh_full<TRosenbrock>(r, nsteps, sv);
    case AMH_TARGET_IID_NORMAL: return launch_mh_d2<TIidNormal>(r, nsteps, sv);
    case AMH_TARGET_NIG_TOY:
    case AMH_TARGET_NIG_TOY_LOG: return launch_mh_d2<TNig>(r, nsteps, sv);
    case AMH_TARGET_LOGISTIC: return launch_mh_t<0, TLogistic>(r, nsteps, sv);
    case AMH_TARGET_USER: return launch_mh_t<0, TUser>(r, nsteps, sv);
    }
    return fail(AMH_ERR_INVALID, "unknown target kind");
}

/* ------------------------------------------------------------------- init */
template <class T>
int launch_init_t(amh_run& r, int mode) {
    const amh_sampler& s = *r.sampler;
    InitArgs a;
    std::memset(&a, 0, sizeof(a));
    a.st = chain_state(r);
    a.d = r.dim;
    a.mode = mode;
    a.want_grad = s.d.kind == AMH_SAMPLER_MALA;
    a.want_lq = (s.d.kind == AMH_SAMPLER_STATIC && !s.d.symmetric);
    a.init_acc = r.keep_acc ? -1 : (s.d.kind == AMH_SAMPLER_RAM) ? 1 : 0;
    a.n_walkers = s.d.n_walkers > 0 ? s.d.n_walkers : 1;
    a.prop = make_prop<0>(s);
    a.S = r.S;
    a.S0 = s.dS0;
    a.comps = s.dcomps;
    if (s.by_components()) a.want_lq = 0;        /* component proposals recompute logq(state) every step */
    const auto tp = make_tp<T, 0>(*r.target);
    const unsigned grid = (unsigned)((r.n + 63) / 64);
    if constexpr (T::kind == AMH_TARGET_USER) {
        void* params[] = {(void*)&a, (void*)&tp};
        const int rc = rtc_launch(r, RK_INIT, grid, 64, 0, params);
        if (rc) return rc;
    } else {
        init_kernel<T><<<grid, 64, 0, r.ctx->stream>>>(a, tp);
        AMH_CUDA_TRY(cudaGetLastError());
    }
    r.launches += 1;
    return AMH_OK;
}

/* static MH keeps logq(state) cached (mh-core.jl:119-123 evaluates it per step); amh_run_set_state recomputes it */
__global__ void __launch_bounds__(64)
relq_kernel(const __grid_constant__ ChainState st, int d, const __grid_constant__ PropP<0> prop) {
    const long long ch = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (ch >= st.n) return;
    double x[Dim<0>::cap];
    for (int i = 0; i < d; ++i) x[i] = st.X[(long long)i * st.pitch + ch];
    st.lq[ch] = logq<0>(x, d, prop);
}

int launch_relq(amh_run& r) {
    const unsigned grid = (unsigned)((r.n + 63) / 64);
    relq_kernel<<<grid, 64, 0, r.ctx->stream>>>(chain_state(r), r.dim, make_prop<0>(*r.sampler));
    AMH_CUDA_TRY(cudaGetLastError());
    r.launches += 1;
    return AMH_OK;
}

int launch_init(amh_run& r, int mode) {
    if (r.dim > kGenericCap) return fail(AMH_ERR_UNSUPPORTED, "device samplers support dim <= 128");
    switch (r.target->kind) {
    case AMH_TARGET_MVNORMAL: return launch_init_t<TMvNormal>(r, mode);
    case AMH_TARGET_GAUSS_PREC: return launch_init_t<TGaussPrec>(r, mode);
    case AMH_TARGET_ROSENBROCK: return launch_init_t<TRosenbrock>(r, mode);
    case AMH_TARGET_IID_NORMAL: return launch_init_t<TIidNormal>(r, mode);
    case AMH_TARGET_NIG_TOY:
    case AMH_TARGET_NIG_TOY_LOG: return launch_init_t<TNig>(r, mode);
    case AMH_TARGET_LOGISTIC: return launch_init_t<TLogistic>(r, mode);
    case AMH_TARGET_USER: return launch_init_t<TUser>(r, mode);
    }
    return fail(AMH_ERR_INVALID, "unknown target kind");
}

#endif  /* AMH_MH_EXTRA_TU */

}  // namespace amhh
