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
 * K2F: same exact sequential semantics, split in two kernels.
 *
 *   stretch_noise_kernel   everything of a launch's sweeps that does not depend on the walker positions -- partner
 *                          index, stretch factor z, (d-1) log z, the exponential -- for ALL sweeps, ensembles and
 *                          walkers at once: embarrassingly parallel, runs on all SMs (an ensemble kernel alone only
 *                          occupies one SM per ensemble).
 *   stretch_sweep_fast_kernel   one CTA per ensemble; per sweep the dependency forest is executed in wavefronts:
 *                          `done[i]` holds the wavefront in which walker i was finished, so "partner available" is
 *                          done[idx] != 0 && done[idx] < wavefront and ONE barrier per wavefront (the
 *                          __syncthreads_or that also detects completion) orders flags and walker data.  Log-densities,
 *                          accept flags and counters of the ensemble live in shared memory for the whole launch, so a
 *                          move costs one round trip to L2 (its 2 x d coordinates + 3 precomputed numbers). */
struct StretchNoise {
    int* partner;              /* [nsteps][n] */
    double* zf;                /* [nsteps][n] */
    double* am;
    double* ex;
};

__global__ void __launch_bounds__(256)
stretch_noise_kernel(StretchNoise o, const unsigned long long* __restrict__ seeds, long long n, int nw, int d, int nsteps,
                     unsigned long long step0, double aa) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n * nsteps) return;
    const int s = (int)(t / n);
    const long long g = t % n;
    const long long en = g / nw;
    const int i = (int)(g % nw);
    const unsigned long long k = step0 + (unsigned long long)s + 1ull;
    const unsigned long long seed = seeds[en];
    const unsigned long long blk = (k * (unsigned long long)nw + (unsigned long long)i) * 2ull;
    const amh::Block b0 = amh::stream_block(seed, blk, 0u);
    const amh::Block b1 = amh::stream_block(seed, blk + 1ull, 0u);
    /* idx = mod1(i + rand(1:(n-1)), n)  (emcee.jl:52) */
    const long long rr = (long long)amh::bounded(b0.v[0], b0.v[1], (unsigned long long)(nw - 1));
    o.partner[t] = (int)((i + rr + 1) % nw);
    const double u = amh::u01(b0.v[2], b0.v[3]);
    const double tt = (aa - 1.0) * u + 1.0;
    const double z = (tt * tt) / aa;
    o.zf[t] = z;
    o.am[t] = (double)(d - 1) * amh::log_(z);
    o.ex[t] = amh::exponential(b1.v[0], b1.v[1]);
}

template <int DMAX, class T, int BLOCK>
__global__ void __launch_bounds__(BLOCK)
stretch_sweep_fast_kernel(const __grid_constant__ StretchArgs a, const __grid_constant__ StretchNoise pre,
                          const __grid_constant__ typename T::template Params<DMAX> tp) {
    using D = Dim<DMAX>;
    constexpr int CAP = D::cap;
    constexpr int UNR = D::unr;
    extern __shared__ __align__(16) double smem_d[];
    const int nw = (int)a.n_walkers;
    double* lpa = smem_d;                                                    /* [nw] log-density, sweep buffer A */
    double* lpb = lpa + nw;                                                  /* [nw]                    buffer B */
    int* partner = reinterpret_cast<int*>(lpb + nw);                         /* [nw]                             */
    int* list_ = partner + nw;                                               /* [nw] work list of a wavefront    */
    unsigned* naccs = reinterpret_cast<unsigned*>(list_ + nw);               /* [nw] accepted moves this launch  */
    unsigned short* done = reinterpret_cast<unsigned short*>(naccs + nw);    /* [nw] wavefront of completion     */
    unsigned char* accs = reinterpret_cast<unsigned char*>(done + nw);       /* [nw] last accept flag            */
    int* list = list_;
    __shared__ int cnt[2];
    const int tid = threadIdx.x;
    const long long en = blockIdx.x;
    const long long base = en * nw;
    const long long n = a.st.n;
    const int d = D::fixed ? DMAX : a.d;
    const int top = D::fixed ? DMAX : d;
    const long long pitch = a.st.pitch;
    double* Xold = a.st.X;  double* Xnew = a.X2;
    double* lpo = lpa;      double* lpn = lpb;
    for (int i = tid; i < nw; i += BLOCK) {
        lpa[i] = a.st.lp[base + i];
        naccs[i] = 0u;
        accs[i] = a.st.acc[base + i];
    }

    for (int s = 0; s < a.nsteps; ++s) {
        const long long off = (long long)s * n + base;
        for (int i = tid; i < nw; i += BLOCK) {
            partner[i] = pre.partner[off + i];
            done[i] = 0;
        }
        if (tid < 2) cnt[tid] = 0;
        __syncthreads();
        unsigned wf = 1;
        int pending;
        const int tmax = (nw + BLOCK - 1) / BLOCK;
        do {
            pending = 0;
            /* (1) compact the walkers whose partner is available into a work list (warp-aggregated append), so that
             *     the moves below run with full warps instead of a few ready lanes per warp */
            int* mycnt = cnt + (wf & 1);
#pragma unroll 1
            for (int t = 0; t < tmax; ++t) {
                const int i = tid + t * BLOCK;
                bool cand = false, ready = false;
                if (i < nw && !done[i]) {
                    cand = true;
                    const int idx = partner[i];
                    const unsigned dn = done[idx];
                    ready = (idx > i) || (dn != 0u && dn < wf);
                }
                const unsigned m = __ballot_sync(0xffffffffu, ready);
                if (m) {
                    const int lane = tid & 31;
                    const int leader = __ffs(m) - 1;
                    int pos = 0;
                    if (lane == leader) pos = atomicAdd(mycnt, __popc(m));
                    pos = __shfl_sync(0xffffffffu, pos, leader);
                    if (ready) list[pos + __popc(m & ((1u << lane) - 1u))] = i;
                }
                if (cand && !ready) pending = 1;
            }
            __syncthreads();
            const int nready = *mycnt;
            if (tid == 0) cnt[(wf + 1) & 1] = 0;                          /* the other counter serves the next wavefront */
            /* (2) the moves of this wavefront */
#pragma unroll 1
            for (int q = tid; q < nready; q += BLOCK) {
                const int i = list[q];
                const int idx = partner[i];
                const double* other = (idx < i) ? Xnew : Xold;          /* emcee.jl:53 */
                const double z = pre.zf[off + i];
                const double am = pre.am[off + i];
                const double ex = pre.ex[off + i];
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
                const double lpw = lpo[i];
                const double alpha = (am + lpy) - lpw;
                const bool acc = (-ex <= alpha);                         /* emcee.jl:93 (non-strict) */
#pragma unroll UNR
                for (int j = 0; j < top; ++j)
                    if (j < d) Xnew[(long long)j * pitch + base + i] = acc ? y[j] : w[j];
                lpn[i] = acc ? lpy : lpw;
                accs[i] = acc ? 1 : 0;
                if (acc) naccs[i] += 1u;
                done[i] = (unsigned short)wf;
            }
            ++wf;
            pending = __syncthreads_or(pending);
        } while (pending);
        double* tX = Xold; Xold = Xnew; Xnew = tX;
        double* tl = lpo; lpo = lpn; lpn = tl;
    }
    /* write the ensemble's scalars back: the current log-densities go to the buffer that pairs with Xold
     * (the host swaps its X / lp pointers when the number of sweeps is odd) */
    double* lpg = (a.nsteps & 1) ? a.lp2 : a.st.lp;
    for (int i = tid; i < nw; i += BLOCK) {
        const long long ch = base + i;
        lpg[ch] = lpo[i];
        a.st.acc[ch] = accs[i];
        a.st.nacc[ch] = a.st.nacc[ch] + (unsigned long long)naccs[i];
        if (a.sv.out || a.sv.sum) {
            for (int j = 0; j < d; ++j) {
                const double v = Xold[(long long)j * pitch + ch];
                if (a.sv.out) a.sv.out[(long long)j * a.sv.out_pitch + ch] = v;
                if (a.sv.sum) {
                    const long long o = (long long)j * pitch + ch;
                    a.sv.sum[o] = a.sv.sum[o] + v;
                    a.sv.sumsq[o] = fma(v, v, a.sv.sumsq[o]);
                }
            }
        }
        if (a.sv.out) a.sv.out[(long long)d * a.sv.out_pitch + ch] = lpo[i];
        if (a.sv.acc_out) a.sv.acc_out[ch] = accs[i];
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
    if (a.n_walkers <= 7000) {          /* 31 bytes of shared memory per walker */
        /* state-independent draws of all `nsteps` sweeps, on all SMs */
        const size_t per = (size_t)r.n * sizeof(double);
        const size_t need = (size_t)nsteps * (3 * per + (size_t)r.n * sizeof(int));
        if (need > r.scratch_bytes) {
            dfree(r.ctx, r.scratch);
            r.scratch = nullptr; r.scratch_bytes = 0;
            const int rca = dmalloc(r.ctx, &r.scratch, need);
            if (rca) return rca;
            r.scratch_bytes = need;
        }
        StretchNoise pre;
        pre.zf = (double*)r.scratch;
        pre.am = pre.zf + (size_t)nsteps * r.n;
        pre.ex = pre.am + (size_t)nsteps * r.n;
        pre.partner = (int*)(pre.ex + (size_t)nsteps * r.n);
        if (nsteps > 0) {
            const long long total = (long long)nsteps * r.n;
            stretch_noise_kernel<<<(unsigned)((total + 255) / 256), 256, 0, r.ctx->stream>>>(pre, r.seeds, r.n, (int)a.n_walkers, r.dim,
                                                                                             nsteps, a.step0, a.a);
            AMH_CUDA_TRY(cudaGetLastError());
            r.launches += 1;
        }
        const size_t smemf = (size_t)a.n_walkers * (2 * sizeof(double) + 2 * sizeof(int) + sizeof(unsigned) + sizeof(unsigned short) + 1) + 16;
        auto kf = stretch_sweep_fast_kernel<DMAX, T, BLOCK>;
        if (smemf > 48 * 1024) AMH_CUDA_TRY(cudaFuncSetAttribute(kf, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemf));
        const unsigned gridf = (unsigned)(r.n / a.n_walkers);
        kf<<<gridf, BLOCK, smemf, r.ctx->stream>>>(a, pre, tp);
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
