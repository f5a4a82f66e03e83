/* amh_launch_ram.cu -- K4: Robust Adaptive Metropolis (Vihola 2012),
 * RobustAdaptiveMetropolis.jl:123-173 (inner step + adaptation) and :216-278
 * (step / step_warmup), with the rank-1 Cholesky up/down-date of Julia's
 * LinearAlgebra fused into the same kernel.
 *
 * Data layout: every chain owns a lower-triangular factor S, stored packed by
 * rows as S[tri(i,j)][chain] -- chains fastest, so each element access of a
 * warp is one contiguous 256-byte request whatever the traversal order.  One
 * thread per chain; the three d-vectors a step needs (x, U -> x_new, S U -> v)
 * live in shared memory as [i][thread].  The kernel is HBM-streaming: a warm-up
 * step reads S twice (S U, then the Givens sweep) and writes it once.
 *
 * S is double-buffered with a per-chain selector: the sweep reads the current
 * buffer and writes the other one, and the selector flips only if the new
 * factor is accepted (eigenvalue bounds, :239-264; failed downdate = the
 * reference's PosDefException, recorded in `failed`).  That is the
 * non-mutating lowrankupdate/lowrankdowndate of the reference (:167,:170).
 */
#include "amh_params.cuh"

namespace amhh {
using namespace amhd;

/* @rtc-begin: the device code from here to @rtc-end is also compiled by NVRTC for user-supplied targets (amh_rtc.cu) */
struct RamArgs {
    ChainState st;
    SaveArgs sv;
    int d;
    int nsteps;
    int warmup;
    unsigned long long step0;
    double* S;                 /* buffer 0 */
    double* S2;                /* buffer 1 */
    unsigned char* sflag;      /* which buffer is current */
    unsigned char* failed;
    double* logalpha;
    double* eta;
    double alpha, gamma, lo, hi;
    int check;
};

template <class T, int BLOCK>
__global__ void __launch_bounds__(BLOCK)
ram_step_kernel(const __grid_constant__ RamArgs a, const __grid_constant__ typename T::template Params<0> tp) {
    constexpr int CAP = Dim<0>::cap;
    extern __shared__ double smem[];
    const int tid = threadIdx.x;
    const long long ch = (long long)blockIdx.x * BLOCK + tid;
    if (ch >= a.st.n) return;
    const int d = a.d;
    const long long pitch = a.st.pitch;
    double* sx = smem + tid;                              /* x        : sx[i*BLOCK] */
    double* su = smem + (size_t)d * BLOCK + tid;          /* U, x_new               */
    double* sv = smem + 2 * (size_t)d * BLOCK + tid;      /* S U, then v            */
    const unsigned long long seed = a.st.seeds[ch];
    double lp = a.st.lp[ch], logalpha = a.logalpha[ch], eta = a.eta[ch];
    unsigned long long nacc = a.st.nacc[ch];
    unsigned char accepted = a.st.acc[ch], failed = a.failed[ch], flag = a.sflag[ch];
    for (int i = 0; i < d; ++i) sx[i * BLOCK] = a.st.X[(long long)i * pitch + ch];
    const int cv = a.st.cv;
    const unsigned long long B = amh::blocks_per_step_cv(cv, d);

    for (int s = 0; s < a.nsteps; ++s) {
        const unsigned long long k = a.step0 + (unsigned long long)s + 1ull;   /* = state.iteration */
        const unsigned long long blk0 = k * B;
        const double* __restrict__ Sc = (flag ? a.S2 : a.S) + ch;
        double* __restrict__ Sn = (flag ? a.S : a.S2) + ch;
        /* U = randn(rng, d)  (:135) */
        if (cv == AMH_CONTRACT_V2) {
            for (int j = 0; 4 * j < d; ++j) {
                const amh::Block b = amh::stream_block7(seed, blk0 + (unsigned long long)j, 0u);
                double q0, q1, q2, q3;
                amh::normal_quad(b, q0, q1, q2, q3);
                su[(4 * j) * BLOCK] = q0;
                if (4 * j + 1 < d) su[(4 * j + 1) * BLOCK] = q1;
                if (4 * j + 2 < d) su[(4 * j + 2) * BLOCK] = q2;
                if (4 * j + 3 < d) su[(4 * j + 3) * BLOCK] = q3;
            }
        } else {
            for (int j = 0; 2 * j < d; ++j) {
                const amh::Block b = amh::stream_block(seed, blk0 + (unsigned long long)j, 0u);
                double z0, z1;
                amh::normal_pair(b, z0, z1);
                su[(2 * j) * BLOCK] = z0;
                if (2 * j + 1 < d) su[(2 * j + 1) * BLOCK] = z1;
            }
        }
        /* S U (row by row); x_new = muladd(S, U, x)  (:136) */
        double nu2 = 0.0;
        for (int i = 0; i < d; ++i) {
            const long long r0 = (long long)tri(i, 0) * pitch;
            double t = Sc[r0] * su[0];
#pragma unroll 8
            for (int j = 1; j <= i; ++j) t = fma(Sc[r0 + (long long)j * pitch], su[j * BLOCK], t);
            sv[i * BLOCK] = t;
        }
        for (int i = 0; i < d; ++i) {
            const double ui = su[i * BLOCK];
            nu2 = (i == 0) ? ui * ui : fma(ui, ui, nu2);
        }
        double xl[CAP];
        for (int i = 0; i < d; ++i) {
            const double xn = sv[i * BLOCK] + sx[i * BLOCK];
            su[i * BLOCK] = xn;                               /* U is dead: keep x_new there */
            xl[i] = xn;
        }
        const double lp_new = T::template logp<0>(xl, d, tp);
        const double dl = lp_new - lp;
        logalpha = (dl != dl) ? dl : (dl < 0.0 ? dl : 0.0);    /* min(lp_new - lp, 0)  (:147) */
        const amh::Block be = amh::step_block(cv, seed, blk0 + (unsigned long long)amh::normal_blocks(cv, d));
        const double e = amh::exponential(be.v[0], be.v[1]);
        const bool isaccept = e > -logalpha;                   /* (:148) */
        if (a.warmup) {
            /* ram_adapt (:153-173) */
            const double dalpha = amh::exp_(logalpha) - a.alpha;
            eta = amh::exp_(-a.gamma * amh::log_((double)(long long)k));      /* iteration^(-gamma) */
            if (dalpha == dalpha) {
                const double cc = sqrt(eta * fabs(dalpha));
                const double nu = sqrt(nu2);
                for (int i = 0; i < d; ++i) sv[i * BLOCK] = (cc * sv[i * BLOCK]) / nu;
                bool ok = true;
                if (dalpha > 0.0) {
                    /* lowrankupdate: Givens sweep over columns */
                    for (int i = 0; i < d; ++i) {
                        const double f = Sc[(long long)tri(i, i) * pitch], g = sv[i * BLOCK];
                        const double rr = sqrt(fma(f, f, g * g));
                        const double c = f / rr, sn = g / rr;
                        Sn[(long long)tri(i, i) * pitch] = rr;
                        if (a.check && !(a.lo <= rr && rr <= a.hi)) ok = false;
#pragma unroll 8
                        for (int j = i + 1; j < d; ++j) {
                            const long long o = (long long)tri(j, i) * pitch;
                            const double Aji = Sc[o];
                            const double vj = sv[j * BLOCK];
                            Sn[o] = c * Aji + sn * vj;
                            sv[j * BLOCK] = c * vj - sn * Aji;
                        }
                    }
                } else {
                    /* lowrankdowndate; s^2 > 1 is the reference's PosDefException */
                    for (int i = 0; i < d && ok; ++i) {
                        const double Aii = Sc[(long long)tri(i, i) * pitch];
                        const double sn = sv[i * BLOCK] / Aii;
                        const double s2 = sn * sn;
                        if (s2 > 1.0) { ok = false; failed = 1; break; }
                        const double c = sqrt(1.0 - s2);
                        const double dg = c * Aii;
                        Sn[(long long)tri(i, i) * pitch] = dg;
                        if (a.check && !(a.lo <= dg && dg <= a.hi)) ok = false;
#pragma unroll 8
                        for (int j = i + 1; j < d; ++j) {
                            const long long o = (long long)tri(j, i) * pitch;
                            const double Aji = (Sc[o] - sn * sv[j * BLOCK]) / c;
                            Sn[o] = Aji;
                            sv[j * BLOCK] = -sn * Aji + c * sv[j * BLOCK];
                        }
                    }
                }
                if (ok) flag ^= 1;        /* S_new becomes current; otherwise S is kept (:259-264) */
            } else {
                failed = 1;
            }
        }
        if (isaccept) {
            for (int i = 0; i < d; ++i) sx[i * BLOCK] = su[i * BLOCK];
            lp = lp_new;
            ++nacc;
        }
        accepted = isaccept ? 1 : 0;
    }

    for (int i = 0; i < d; ++i) {
        const double v = sx[i * BLOCK];
        const long long o = (long long)i * pitch + ch;
        a.st.X[o] = v;
        if (a.sv.out) a.sv.out[(long long)i * a.sv.out_pitch + ch] = v;
        if (a.sv.sum) {
            save_moments(a.sv, o, v);
        }
    }
    a.st.lp[ch] = lp;
    a.st.nacc[ch] = nacc;
    a.st.acc[ch] = accepted;
    a.logalpha[ch] = logalpha;
    a.eta[ch] = eta;
    a.failed[ch] = failed;
    a.sflag[ch] = flag;
    if (a.sv.out) a.sv.out[(long long)d * a.sv.out_pitch + ch] = lp;
    if (a.sv.acc_out) a.sv.acc_out[ch] = accepted;
}

/* @rtc-end */
/* gathers the current buffer of every chain into `dst` ([tri][pitch]) for get_state */
__global__ void ram_gather_S_kernel(const double* S, const double* S2, const unsigned char* sflag, double* dst,
                                    long long n, long long pitch, long long nt) {
    const long long ch = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (ch >= n) return;
    const double* src = sflag[ch] ? S2 : S;
    for (long long q = 0; q < nt; ++q) dst[q * pitch + ch] = src[q * pitch + ch];
}

template <class T>
int launch_ram_t(amh_run& r, int nsteps, bool warmup, const SaveArgs& sv) {
    const amh_sampler& s = *r.sampler;
    const int d = r.dim;
    /* 3 vectors of d doubles per thread in shared memory */
    const size_t per_thread = 3 * (size_t)d * sizeof(double);
    int block = 128;
    while (block > 32 && per_thread * block > 200 * 1024) block >>= 1;
    RamArgs a;
    std::memset(&a, 0, sizeof(a));
    a.st = chain_state(r);
    a.sv = sv;
    a.d = d;
    a.nsteps = nsteps;
    a.warmup = warmup ? 1 : 0;
    a.step0 = (unsigned long long)r.step;
    a.S = r.S; a.S2 = r.S2; a.sflag = r.sflag; a.failed = r.failed;
    a.logalpha = r.logalpha; a.eta = r.eta;
    a.alpha = s.d.ram_alpha; a.gamma = s.d.ram_gamma; a.lo = s.d.ram_eig_lo; a.hi = s.d.ram_eig_hi;
    a.check = !(a.lo == 0.0 && a.hi == INFINITY);
    const auto tp = make_tp<T, 0>(*r.target);
    const size_t smem = per_thread * block;
    const unsigned grid = (unsigned)((r.n + block - 1) / block);
    if constexpr (T::kind == AMH_TARGET_USER) {
        void* params[] = {(void*)&a, (void*)&tp};
        const int rc = rtc_launch(r, block == 128 ? RK_RAM128 : block == 64 ? RK_RAM64 : RK_RAM32, grid, (unsigned)block, smem, params);
        if (rc) return rc;
    } else {
#define AMH_RAM_LAUNCH(BL)                                                                                             \
    do {                                                                                                               \
        auto kern = ram_step_kernel<T, BL>;                                                                            \
        if (smem > 48 * 1024)                                                                                          \
            AMH_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));          \
        kern<<<grid, BL, smem, r.ctx->stream>>>(a, tp);                                                                \
    } while (0)
        if (block == 128) AMH_RAM_LAUNCH(128);
        else if (block == 64) AMH_RAM_LAUNCH(64);
        else AMH_RAM_LAUNCH(32);
#undef AMH_RAM_LAUNCH
        AMH_CUDA_TRY(cudaGetLastError());
    }
    r.launches += 1;
    r.pending_launches += 1;
    return AMH_OK;
}

int launch_ram(amh_run& r, int nsteps, bool warmup, const SaveArgs& sv) {
    switch (r.target->kind) {
    case AMH_TARGET_MVNORMAL: return launch_ram_t<TMvNormal>(r, nsteps, warmup, sv);
    case AMH_TARGET_GAUSS_PREC: return launch_ram_t<TGaussPrec>(r, nsteps, warmup, sv);
    case AMH_TARGET_ROSENBROCK: return launch_ram_t<TRosenbrock>(r, nsteps, warmup, sv);
    case AMH_TARGET_IID_NORMAL: return launch_ram_t<TIidNormal>(r, nsteps, warmup, sv);
    case AMH_TARGET_NIG_TOY:
    case AMH_TARGET_NIG_TOY_LOG: return launch_ram_t<TNig>(r, nsteps, warmup, sv);
    case AMH_TARGET_LOGISTIC: return launch_ram_t<TLogistic>(r, nsteps, warmup, sv);
    case AMH_TARGET_USER: return launch_ram_t<TUser>(r, nsteps, warmup, sv);
    }
    return fail(AMH_ERR_INVALID, "unknown target kind");
}

/* set_state: `src` ([tri][pitch]) becomes the current factor of every chain (buffer 0) */
int ram_scatter_S(amh_run& r, const double* src) {
    const size_t nt = (size_t)r.dim * (r.dim + 1) / 2;
    AMH_CUDA_TRY(cudaMemcpyAsync(r.S, src, sizeof(double) * nt * r.pitch, cudaMemcpyDeviceToDevice, r.ctx->stream));
    AMH_CUDA_TRY(cudaMemsetAsync(r.sflag, 0, (size_t)r.pitch, r.ctx->stream));
    return AMH_OK;
}

int ram_gather_S(amh_run& r, double* dst) {
    const long long nt = (long long)r.dim * (r.dim + 1) / 2;
    const unsigned grid = (unsigned)((r.n + 127) / 128);
    ram_gather_S_kernel<<<grid, 128, 0, r.ctx->stream>>>(r.S, r.S2, r.sflag, dst, r.n, r.pitch, nt);
    AMH_CUDA_TRY(cudaGetLastError());
    r.launches += 1;
    return AMH_OK;
}

}  // namespace amhh
