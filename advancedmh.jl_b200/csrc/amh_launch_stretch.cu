/* amh_launch_stretch.cu -- K2: affine-invariant ensemble sampler, stretch move
 * (emcee.jl:39-58 sweep, :70-102 move).
 *
 * The reference sweep is SEQUENTIAL (Gauss-Seidel): walker i's partner idx is
 * read from new_walkers when idx < i and from the old walkers otherwise
 * (emcee.jl:53).  All random draws of a sweep are state independent (partner
 * index, stretch factor, exponential), so the only true dependency is
 * "walker i needs the NEW value of walker idx_i < i".  That dependency graph
 * is a random recursive forest of depth O(log n): the kernel executes it in
 * wavefronts and reproduces the sequential result EXACTLY.
 *
 * One CTA per ensemble.  Walker states are double-buffered in global memory
 * (old sweep / new sweep, [dim][chain] layout, L2 resident); partner indices
 * and "done" flags live in shared memory.  A move is executed by the thread
 * that owns the walker as soon as its partner is available; __syncthreads()
 * between wavefronts orders the global writes inside the CTA.
 */
#include "amh_params.cuh"

namespace amhh {
using namespace amhd;

struct StretchArgs {
    ChainState st;              /* X/lp = buffer A */
    SaveArgs sv;
    double* X2;                 /* buffer B */
    double* lp2;
    int d;
    int nsteps;
    unsigned long long step0;
    long long n_walkers;
    double a;                   /* stretch_length */
};

template <int DMAX, class T, int BLOCK>
__global__ void __launch_bounds__(BLOCK)
stretch_sweep_kernel(const __grid_constant__ StretchArgs a, const __grid_constant__ typename T::template Params<DMAX> tp) {
    using D = Dim<DMAX>;
    constexpr int CAP = D::cap;
    constexpr int UNR = D::unr;
    extern __shared__ int smem_i[];
    const long long nw = a.n_walkers;
    int* partner = smem_i;                                        /* [nw] */
    unsigned char* done = (unsigned char*)(smem_i + nw);          /* [nw] */
    const int tid = threadIdx.x;
    const long long en = blockIdx.x;
    const long long base = en * nw;
    const int d = D::fixed ? DMAX : a.d;
    const int top = D::fixed ? DMAX : d;
    const long long pitch = a.st.pitch;
    const unsigned long long seed = a.st.seeds[en];
    double* Xold = a.st.X;  double* lpold = a.st.lp;
    double* Xnew = a.X2;    double* lpnew = a.lp2;

    for (int s = 0; s < a.nsteps; ++s) {
        const unsigned long long k = a.step0 + (unsigned long long)s + 1ull;
        /* phase 0: partner indices of the whole sweep (state independent) */
        for (long long i = tid; i < nw; i += BLOCK) {
            const unsigned long long blk = (k * (unsigned long long)nw + (unsigned long long)i) * 2ull;
            const amh::Block b0 = amh::stream_block(seed, blk, 0u);
            /* idx = mod1(i + rand(1:(n-1)), n)  (emcee.jl:52) */
            const long long rr = (long long)amh::bounded(b0.v[0], b0.v[1], (unsigned long long)(nw - 1));
            partner[i] = (int)((i + rr + 1) % nw);
            done[i] = 0;
        }
        __syncthreads();
        /* wavefronts */
        int pending;
        do {
            pending = 0;
            /* decide with the flags of the PREVIOUS wavefront, then publish after the barrier */
            for (long long i = tid; i < nw; i += BLOCK) {
                if (done[i]) continue;
                const int idx = partner[i];
                const bool ready = (idx > i) || (done[idx] == 1);
                if (!ready) { pending = 1; continue; }
                const unsigned long long blk = (k * (unsigned long long)nw + (unsigned long long)i) * 2ull;
                const amh::Block b0 = amh::stream_block(seed, blk, 0u);
                const amh::Block b1 = amh::stream_block(seed, blk + 1ull, 0u);
                const double* other = (idx < i) ? Xnew : Xold;      /* emcee.jl:53 */
                /* move (emcee.jl:70-102) */
                const double u = amh::u01(b0.v[2], b0.v[3]);
                const double t = (a.a - 1.0) * u + 1.0;
                const double z = (t * t) / a.a;
                const double alphamult = (double)(d - 1) * amh::log_(z);
                double y[CAP], w[CAP];
#pragma unroll UNR
                for (int j = 0; j < top; ++j)
                    if (j < d) {
                        const double wj = Xold[(long long)j * pitch + base + i];
                        const double oj = other[(long long)j * pitch + base + idx];
                        w[j] = wj;
                        y[j] = oj + z * (wj - oj);
                    }
                const double lpy = T::template logp<DMAX>(y, d, tp);
                const double lpw = lpold[base + i];
                const double alpha = (alphamult + lpy) - lpw;
                const double e = amh::exponential(b1.v[0], b1.v[1]);
                const bool acc = (-e <= alpha);                      /* emcee.jl:93 (non-strict) */
#pragma unroll UNR
                for (int j = 0; j < top; ++j)
                    if (j < d) Xnew[(long long)j * pitch + base + i] = acc ? y[j] : w[j];
                lpnew[base + i] = acc ? lpy : lpw;
                a.st.acc[base + i] = acc ? 1 : 0;
                if (acc) a.st.nacc[base + i] += 1ull;
                done[i] = 2;                                          /* finished in THIS wavefront */
            }
            __syncthreads();
            for (long long i = tid; i < nw; i += BLOCK)
                if (done[i] == 2) done[i] = 1;
            pending = __syncthreads_or(pending);
        } while (pending);
        /* swap buffers */
        double* tX = Xold; Xold = Xnew; Xnew = tX;
        double* tl = lpold; lpold = lpnew; lpnew = tl;
    }
    /* epilogue: Xold/lpold hold the current state (the host swaps its pointers when nsteps is odd) */
    if (a.sv.out || a.sv.sum || a.sv.acc_out) {
        for (long long i = tid; i < nw; i += BLOCK) {
            const long long ch = base + i;
            for (int j = 0; j < d; ++j) {
                const double v = Xold[(long long)j * pitch + ch];
                if (a.sv.out) a.sv.out[(long long)j * a.sv.out_pitch + ch] = v;
                if (a.sv.sum) {
                    const long long o = (long long)j * pitch + ch;
                    a.sv.sum[o] = a.sv.sum[o] + v;
                    a.sv.sumsq[o] = fma(v, v, a.sv.sumsq[o]);
                }
            }
            if (a.sv.out) a.sv.out[(long long)d * a.sv.out_pitch + ch] = lpold[ch];
            if (a.sv.acc_out) a.sv.acc_out[ch] = a.st.acc[ch];
        }
    }
}

/* ---------------------------------------------------------------------------
 * K2F: same exact sequential semantics, for ensembles of up to WPT * BLOCK walkers.
 *   - everything that does not depend on the walker positions (partner index, stretch factor z, (d-1) log z, the
 *     exponential) is computed for the whole sweep in one parallel phase and kept in registers;
 *   - `done[i]` stores the wavefront in which walker i was finished, so "partner available" is
 *     done[idx] != 0 && done[idx] < wavefront and ONE barrier per wavefront (the __syncthreads_or that also
 *     detects completion) orders both the flags and the walker data. */
template <int DMAX, class T, int BLOCK, int WPT>
__global__ void __launch_bounds__(BLOCK)
stretch_sweep_fast_kernel(const __grid_constant__ StretchArgs a, const __grid_constant__ typename T::template Params<DMAX> tp) {
    using D = Dim<DMAX>;
    constexpr int CAP = D::cap;
    constexpr int UNR = D::unr;
    extern __shared__ __align__(16) double smem_d[];
    const int nw = (int)a.n_walkers;
    double* zf = smem_d;                                                     /* [nw] stretch factor z          */
    double* am = zf + nw;                                                    /* [nw] (d-1) log z               */
    double* ex = am + nw;                                                    /* [nw] exponential draw          */
    int* partner = reinterpret_cast<int*>(ex + nw);                          /* [nw]                           */
    unsigned short* done = reinterpret_cast<unsigned short*>(partner + nw);  /* [nw] wavefront of completion   */
    const int tid = threadIdx.x;
    const long long en = blockIdx.x;
    const long long base = en * nw;
    const int d = D::fixed ? DMAX : a.d;
    const int top = D::fixed ? DMAX : d;
    const long long pitch = a.st.pitch;
    const unsigned long long seed = a.st.seeds[en];
    double* Xold = a.st.X;  double* lpold = a.st.lp;
    double* Xnew = a.X2;    double* lpnew = a.lp2;

    for (int s = 0; s < a.nsteps; ++s) {
        const unsigned long long k = a.step0 + (unsigned long long)s + 1ull;
        for (int i = tid; i < nw; i += BLOCK) {
            const unsigned long long blk = (k * (unsigned long long)nw + (unsigned long long)i) * 2ull;
            const amh::Block b0 = amh::stream_block(seed, blk, 0u);
            const amh::Block b1 = amh::stream_block(seed, blk + 1ull, 0u);
            /* idx = mod1(i + rand(1:(n-1)), n)  (emcee.jl:52) */
            const long long rr = (long long)amh::bounded(b0.v[0], b0.v[1], (unsigned long long)(nw - 1));
            partner[i] = (int)((i + rr + 1) % nw);
            const double u = amh::u01(b0.v[2], b0.v[3]);
            const double tt = (a.a - 1.0) * u + 1.0;
            const double z = (tt * tt) / a.a;
            zf[i] = z;
            am[i] = (double)(d - 1) * amh::log_(z);
            ex[i] = amh::exponential(b1.v[0], b1.v[1]);
            done[i] = 0;
        }
        __syncthreads();
        unsigned wf = 1;
        int pending;
        do {
            pending = 0;
#pragma unroll 1
            for (int i = tid; i < nw; i += BLOCK) {
                if (done[i]) continue;
                const int idx = partner[i];
                const unsigned dn = done[idx];
                const bool ready = (idx > i) || (dn != 0u && dn < wf);
                if (!ready) { pending = 1; continue; }
                const double* other = (idx < i) ? Xnew : Xold;          /* emcee.jl:53 */
                const double z = zf[i];
                double y[CAP], w[CAP];
#pragma unroll UNR
                for (int j = 0; j < top; ++j)
                    if (j < d) {
                        const double wj = Xold[(long long)j * pitch + base + i];
                        const double oj = other[(long long)j * pitch + base + idx];
                        w[j] = wj;
                        y[j] = oj + z * (wj - oj);
                    }
                const double lpy = T::template logp<DMAX>(y, d, tp);
                const double lpw = lpold[base + i];
                const double alpha = (am[i] + lpy) - lpw;
                const bool acc = (-ex[i] <= alpha);                      /* emcee.jl:93 (non-strict) */
#pragma unroll UNR
                for (int j = 0; j < top; ++j)
                    if (j < d) Xnew[(long long)j * pitch + base + i] = acc ? y[j] : w[j];
                lpnew[base + i] = acc ? lpy : lpw;
                a.st.acc[base + i] = acc ? 1 : 0;
                if (acc) a.st.nacc[base + i] += 1ull;
                done[i] = (unsigned short)wf;
            }
            ++wf;
            pending = __syncthreads_or(pending);
        } while (pending);
        double* tX = Xold; Xold = Xnew; Xnew = tX;
        double* tl = lpold; lpold = lpnew; lpnew = tl;
    }
    if (a.sv.out || a.sv.sum || a.sv.acc_out) {
        for (int i = tid; i < nw; i += BLOCK) {
            const long long ch = base + i;
            for (int j = 0; j < d; ++j) {
                const double v = Xold[(long long)j * pitch + ch];
                if (a.sv.out) a.sv.out[(long long)j * a.sv.out_pitch + ch] = v;
                if (a.sv.sum) {
                    const long long o = (long long)j * pitch + ch;
                    a.sv.sum[o] = a.sv.sum[o] + v;
                    a.sv.sumsq[o] = fma(v, v, a.sv.sumsq[o]);
                }
            }
            if (a.sv.out) a.sv.out[(long long)d * a.sv.out_pitch + ch] = lpold[ch];
            if (a.sv.acc_out) a.sv.acc_out[ch] = a.st.acc[ch];
        }
    }
}

template <int DMAX, class T>
int launch_stretch_t(amh_run& r, int nsteps, const SaveArgs& sv) {
    constexpr int BLOCK = 1024;
    const amh_sampler& s = *r.sampler;
    StretchArgs a;
    std::memset(&a, 0, sizeof(a));
    a.st = chain_state(r);
    a.sv = sv;
    a.X2 = r.X2; a.lp2 = r.lp2;
    a.d = r.dim;
    a.nsteps = nsteps;
    a.step0 = (unsigned long long)r.step;
    a.n_walkers = s.d.n_walkers;
    a.a = s.d.stretch_a;
    const auto tp = make_tp<T, DMAX>(*r.target);
    if (a.n_walkers <= 7000) {          /* 30 bytes of shared memory per walker */
        const size_t smemf = (size_t)a.n_walkers * (3 * sizeof(double) + sizeof(int) + sizeof(unsigned short)) + 16;
        auto kf = stretch_sweep_fast_kernel<DMAX, T, BLOCK, 4>;
        if (smemf > 48 * 1024) AMH_CUDA_TRY(cudaFuncSetAttribute(kf, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemf));
        const unsigned gridf = (unsigned)(r.n / a.n_walkers);
        kf<<<gridf, BLOCK, smemf, r.ctx->stream>>>(a, tp);
        AMH_CUDA_TRY(cudaGetLastError());
        if (nsteps & 1) {
            std::swap(r.X, r.X2);
            std::swap(r.lp, r.lp2);
        }
        r.launches += 1;
        r.pending_launches += 1;
        return AMH_OK;
    }
    const size_t smem = (size_t)a.n_walkers * (sizeof(int) + 1) + 16;
    auto kern = stretch_sweep_kernel<DMAX, T, BLOCK>;
    if (smem > 200 * 1024) return fail(AMH_ERR_UNSUPPORTED, "Ensemble on the device supports n_walkers <= 40000");
    if (smem > 48 * 1024) AMH_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const unsigned grid = (unsigned)(r.n / a.n_walkers);
    kern<<<grid, BLOCK, smem, r.ctx->stream>>>(a, tp);
    AMH_CUDA_TRY(cudaGetLastError());
    if (nsteps & 1) {          /* the current state now lives in the other buffer */
        std::swap(r.X, r.X2);
        std::swap(r.lp, r.lp2);
    }
    r.launches += 1;
    r.pending_launches += 1;
    return AMH_OK;
}

template <class T>
int launch_stretch_dim(amh_run& r, int nsteps, const SaveArgs& sv) {
    switch (r.dim) {          /* exact-dimension instantiations; everything else is generic */
    case 2: return launch_stretch_t<2, T>(r, nsteps, sv);
    case 3: return launch_stretch_t<3, T>(r, nsteps, sv);
    case 4: return launch_stretch_t<4, T>(r, nsteps, sv);
    case 5: return launch_stretch_t<5, T>(r, nsteps, sv);
    case 8: return launch_stretch_t<8, T>(r, nsteps, sv);
    case 10: return launch_stretch_t<10, T>(r, nsteps, sv);
    case 16: return launch_stretch_t<16, T>(r, nsteps, sv);
    }
    return launch_stretch_t<0, T>(r, nsteps, sv);
}

int launch_stretch(amh_run& r, int nsteps, const SaveArgs& sv) {
    switch (r.target->kind) {
    case AMH_TARGET_MVNORMAL: return launch_stretch_dim<TMvNormal>(r, nsteps, sv);
    case AMH_TARGET_GAUSS_PREC: return launch_stretch_dim<TGaussPrec>(r, nsteps, sv);
    case AMH_TARGET_ROSENBROCK: return launch_stretch_dim<TRosenbrock>(r, nsteps, sv);
    case AMH_TARGET_IID_NORMAL: return launch_stretch_t<2, TIidNormal>(r, nsteps, sv);
    case AMH_TARGET_NIG_TOY:
    case AMH_TARGET_NIG_TOY_LOG: return launch_stretch_t<2, TNig>(r, nsteps, sv);
    case AMH_TARGET_LOGISTIC: return launch_stretch_t<0, TLogistic>(r, nsteps, sv);
    }
    return fail(AMH_ERR_INVALID, "unknown target kind");
}

}  // namespace amhh
