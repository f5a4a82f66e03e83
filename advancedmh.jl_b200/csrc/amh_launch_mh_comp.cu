/* amh_launch_mh_comp.cu -- K1C: the Metropolis-Hastings step for proposals that are ARRAYS of univariate laws.
 *
 *   one Proposal over an array of distributions   StaticProposal([Normal(0,1), InverseGamma(2,3)])   README.md:106
 *       rand = map(rand, p.proposal), logpdf = left-to-right sum               proposal.jl:26-35, 41-85
 *   an array of Proposals (AMH_SAMPLER_MIXED)     [StaticProposal(Normal(0,1)), RandomWalkProposal(Normal(0,.1))]
 *       per-coordinate static / random walk, own `issymmetric`                 proposal.jl:132-150, 195-196, 236-240
 *
 * One thread per chain, chains fastest in memory like every other kernel; the state sits in shared memory
 * ([i][thread], conflict free), candidate and scratch vectors in local memory (runtime dimension <= 128).  The
 * families with a rejection sampler (Gamma / InverseGamma) read private Philox sub-streams, so lanes that need
 * another attempt do not shift anyone's stream (include/amh_contract.h). */
#include "amh_params.cuh"

namespace amhh {
using namespace amhd;

/* @rtc-begin: the device code from here to @rtc-end is also compiled by NVRTC for user-supplied targets (amh_rtc.cu) */
struct MhCompArgs {
    ChainState st;
    SaveArgs sv;
    int d;
    int mixed;                   /* AMH_SAMPLER_MIXED: per-component rw / symmetric flags */
    int is_rw;                   /* otherwise: the proposal's kind ...                    */
    int sym;                     /* ... and its issymmetric                               */
    int nsteps;
    unsigned long long step0;
    const amh_component* comps;
};

template <class T, int BLOCK>
__global__ void __launch_bounds__(BLOCK)
mh_comp_kernel(const __grid_constant__ MhCompArgs a, const __grid_constant__ typename T::template Params<0> tp) {
    constexpr int CAP = Dim<0>::cap;
    extern __shared__ double sx[];
    const int tid = threadIdx.x;
    const long long ch = (long long)blockIdx.x * BLOCK + tid;
    if (ch >= a.st.n) return;
    const int d = a.d;
    const unsigned long long seed = a.st.seeds[ch];
    double lp = a.st.lp[ch];
    unsigned long long nacc = a.st.nacc[ch];
    unsigned char accepted = a.st.acc[ch];
    for (int i = 0; i < d; ++i) sx[i * BLOCK + tid] = a.st.X[(long long)i * a.st.pitch + ch];

    const int cv = a.st.cv;
    const unsigned long long B = amh::blocks_per_step_cv(cv, d);
    double z[CAP];
    for (int s = 0; s < a.nsteps; ++s) {
        const unsigned long long k = a.step0 + (unsigned long long)s + 1ull;
        const unsigned long long blk0 = k * B;
        step_normals<0>(cv, seed, blk0, d, z);
        const amh::Block be = amh::step_block(cv, seed, blk0 + (unsigned long long)amh::normal_blocks(cv, d));
        const double e = amh::exponential(be.v[0], be.v[1]);
        draw_components(z, d, a.comps, seed, k * (unsigned long long)d);
        for (int i = 0; i < d; ++i) {
            const bool rw_i = a.mixed ? (a.comps[i].rw != 0) : (a.is_rw != 0);
            if (rw_i) z[i] = sx[i * BLOCK + tid] + z[i];
        }
        const double lp_c = T::template logp<0>(z, d, tp);
        double logratio = 0.0;
        if (a.mixed) {
            for (int i = 0; i < d; ++i) {
                const amh_component q = a.comps[i];
                const double xi = sx[i * BLOCK + tid];
                double lr = 0.0;
                if (!q.symmetric) {
                    if (q.rw)
                        lr = amh::family_logpdf(q.family, q.p0, q.p1, q.logc, xi - z[i]) -
                             amh::family_logpdf(q.family, q.p0, q.p1, q.logc, z[i] - xi);
                    else
                        lr = amh::family_logpdf(q.family, q.p0, q.p1, q.logc, xi) -
                             amh::family_logpdf(q.family, q.p0, q.p1, q.logc, z[i]);
                }
                logratio = (i == 0) ? lr : logratio + lr;
            }
        } else if (!a.sym) {
            double t1[CAP];
            if (a.is_rw) {
                double t2[CAP];
                for (int i = 0; i < d; ++i) {
                    const double xi = sx[i * BLOCK + tid];
                    t1[i] = xi - z[i];
                    t2[i] = z[i] - xi;
                }
                logratio = logq_components(t1, d, a.comps) - logq_components(t2, d, a.comps);
            } else {
                for (int i = 0; i < d; ++i) t1[i] = sx[i * BLOCK + tid];
                logratio = logq_components(t1, d, a.comps) - logq_components(z, d, a.comps);
            }
        }
        const double loga = (lp_c - lp) + logratio;
        if (-e < loga) {                                   /* mh-core.jl:108 (strict; NaN rejects) */
            for (int i = 0; i < d; ++i) sx[i * BLOCK + tid] = z[i];
            lp = lp_c;
            accepted = 1;
            ++nacc;
        } else {
            accepted = 0;
        }
    }

    for (int i = 0; i < d; ++i) {
        const double v = sx[i * BLOCK + tid];
        const long long o = (long long)i * a.st.pitch + ch;
        a.st.X[o] = v;
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
template <class T>
static int launch_mh_comp_t(amh_run& r, int nsteps, const SaveArgs& sv) {
    constexpr int BLOCK = 64;
    const amh_sampler& s = *r.sampler;
    MhCompArgs a;
    std::memset(&a, 0, sizeof(a));
    a.st = chain_state(r);
    a.sv = sv;
    a.d = r.dim;
    a.mixed = s.d.kind == AMH_SAMPLER_MIXED;
    a.is_rw = s.d.kind == AMH_SAMPLER_RW;
    a.sym = s.d.symmetric != 0;
    a.nsteps = nsteps;
    a.step0 = (unsigned long long)r.step;
    a.comps = s.dcomps;
    const auto tp = make_tp<T, 0>(*r.target);
    const size_t smem = (size_t)r.dim * BLOCK * sizeof(double);
    const unsigned grid = (unsigned)((r.n + BLOCK - 1) / BLOCK);
    if constexpr (T::kind == AMH_TARGET_USER) {
        static_assert(BLOCK == 64, "RK_COMP names mh_comp_kernel<TUser, 64>");
        void* params[] = {(void*)&a, (void*)&tp};
        const int rc = rtc_launch(r, RK_COMP, grid, BLOCK, smem, params);
        if (rc) return rc;
    } else {
        auto kern = mh_comp_kernel<T, BLOCK>;
        if (smem > 48 * 1024) AMH_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<grid, BLOCK, smem, r.ctx->stream>>>(a, tp);
        AMH_CUDA_TRY(cudaGetLastError());
    }
    r.launches += 1;
    r.pending_launches += 1;
    return AMH_OK;
}

int launch_mh_comp(amh_run& r, int nsteps, const SaveArgs& sv) {
    if (r.dim > kGenericCap) return fail(AMH_ERR_UNSUPPORTED, "component proposals on the device support dim <= 128");
    switch (r.target->kind) {
    case AMH_TARGET_MVNORMAL: return launch_mh_comp_t<TMvNormal>(r, nsteps, sv);
    case AMH_TARGET_GAUSS_PREC: return launch_mh_comp_t<TGaussPrec>(r, nsteps, sv);
    case AMH_TARGET_ROSENBROCK: return launch_mh_comp_t<TRosenbrock>(r, nsteps, sv);
    case AMH_TARGET_IID_NORMAL: return launch_mh_comp_t<TIidNormal>(r, nsteps, sv);
    case AMH_TARGET_NIG_TOY:
    case AMH_TARGET_NIG_TOY_LOG: return launch_mh_comp_t<TNig>(r, nsteps, sv);
    case AMH_TARGET_LOGISTIC: return launch_mh_comp_t<TLogistic>(r, nsteps, sv);
    case AMH_TARGET_USER: return launch_mh_comp_t<TUser>(r, nsteps, sv);
    }
    return fail(AMH_ERR_INVALID, "unknown target kind");
}

}  // namespace amhh
