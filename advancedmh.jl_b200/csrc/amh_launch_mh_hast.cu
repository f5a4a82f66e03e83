/* amh_launch_mh_hast.cu -- the per-thread MH step kernel K1 for StaticProposal with issymmetric = false in exact dimensions.
 *
 * `StaticMH(d)` of the reference builds a StaticProposal{false} (proposal.jl:1-21, mh-core.jl:48-51): every step evaluates
 * the Hastings term logq(state) - logq(candidate) (mh-core.jl:119-123, proposal.jl:79-85,190-192).  Until round 2 only the
 * generic kernel (run-time dimension, vectors in local memory) had that term, so the reference's plain StaticMH ran
 * 7-21 x slower than RWMH on the same target: 65 536 chains on a MvNormal target, d = 8 / 16 / 32: 4.1e9 / 1.6e9 / 2.9e8
 * chain-steps/s (profiles/r2_dim_cliffs.txt).  Here the same kernel template is instantiated with the term (HAST1,
 * amh_kernels.cuh: logq_fixed = the generic logq's operations in the same order, vectors in registers).  A translation
 * unit of its own so that it compiles in parallel with the others. */
#define AMH_MH_EXTRA_TU
#include "amh_launch_mh.cu"

namespace amhh {

template <int DMAX, class T>
static int launch_hast_t(amh_run& r, int nsteps, const SaveArgs& sv) {
    constexpr int BLOCK = LaunchCfg<DMAX>::block;
    constexpr int MINB = LaunchCfg<DMAX>::minb;
    const amh_sampler& s = *r.sampler;
    MhArgs<DMAX> a;
    std::memset(&a, 0, sizeof(a));
    a.st = chain_state(r);
    a.sv = sv;
    a.d = r.dim;
    a.is_rw = 0;
    a.hast = 1;
    a.nsteps = nsteps;
    a.step0 = (unsigned long long)r.step;
    a.prop = make_prop<DMAX>(s);
    const auto tp = make_tp<T, DMAX>(*r.target);
    const size_t smem = (size_t)r.dim * BLOCK * sizeof(double);
    const unsigned grid = (unsigned)((r.n + BLOCK - 1) / BLOCK);
    auto kern = mh_step_kernel<DMAX, T, BLOCK, MINB, true>;
    if (smem > 48 * 1024) AMH_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, BLOCK, smem, r.ctx->stream>>>(a, tp);
    AMH_CUDA_TRY(cudaGetLastError());
    r.launches += 1;
    r.pending_launches += 1;
    return AMH_OK;
}

template <class T>
static int hast_dims(amh_run& r, int nsteps, const SaveArgs& sv, bool& taken) {
    taken = true;
    switch (r.dim) {
    case 1: return launch_hast_t<1, T>(r, nsteps, sv);
    case 2: return launch_hast_t<2, T>(r, nsteps, sv);
    case 3: return launch_hast_t<3, T>(r, nsteps, sv);
    case 4: return launch_hast_t<4, T>(r, nsteps, sv);
    case 5: return launch_hast_t<5, T>(r, nsteps, sv);
    case 6: return launch_hast_t<6, T>(r, nsteps, sv);
    case 7: return launch_hast_t<7, T>(r, nsteps, sv);
    case 8: return launch_hast_t<8, T>(r, nsteps, sv);
    case 9: return launch_hast_t<9, T>(r, nsteps, sv);
    case 10: return launch_hast_t<10, T>(r, nsteps, sv);
    case 12: return launch_hast_t<12, T>(r, nsteps, sv);
    case 14: return launch_hast_t<14, T>(r, nsteps, sv);
    case 16: return launch_hast_t<16, T>(r, nsteps, sv);
    case 20: return launch_hast_t<20, T>(r, nsteps, sv);
    case 24: return launch_hast_t<24, T>(r, nsteps, sv);
    case 28: return launch_hast_t<28, T>(r, nsteps, sv);
    case 32: return launch_hast_t<32, T>(r, nsteps, sv);
    }
    taken = false;
    return AMH_OK;
}

int launch_mh_hast(amh_run& r, int nsteps, const SaveArgs& sv, bool& taken) {
    taken = false;
    if (std::getenv("AMH_MH_NO_HAST1")) return AMH_OK;                        /* A/B switch: the generic kernel */
    switch (r.target->kind) {
    case AMH_TARGET_MVNORMAL: return hast_dims<TMvNormal>(r, nsteps, sv, taken);
    case AMH_TARGET_GAUSS_PREC: return hast_dims<TGaussPrec>(r, nsteps, sv, taken);
    case AMH_TARGET_ROSENBROCK: return hast_dims<TRosenbrock>(r, nsteps, sv, taken);
    case AMH_TARGET_IID_NORMAL:
        if (r.dim == 2) { taken = true; return launch_hast_t<2, TIidNormal>(r, nsteps, sv); }
        break;
    case AMH_TARGET_NIG_TOY:
    case AMH_TARGET_NIG_TOY_LOG:
        if (r.dim == 2) { taken = true; return launch_hast_t<2, TNig>(r, nsteps, sv); }
        break;
    }
    return AMH_OK;
}

}  // namespace amhh
