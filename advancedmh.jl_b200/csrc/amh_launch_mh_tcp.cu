/* amh_launch_mh_tcp.cu -- K1T16 for the dimensions in between: the random-walk / static MH step on a MvNormal target
 * (mh-core.jl:92-117, proposal.jl:41-56) with the dimension PADDED to the next multiple of 8, up to 64.
 *
 * Why: the exact-dimension kernels cover d = 1..6, 8, 10, 12, 16, 20, 24, 32; every other dimension fell to the generic
 * per-thread kernel (run-time dimension, vectors in local memory), which is 4-100 x slower -- measured on 65 536 chains
 * (profiles/r2_c2_dims.txt, before -> after): d = 7 -> 7.1e9 chain-steps/s next to 3.0e10 at d = 8 (now 2.8e10), d = 14 -> 2.7e9 (now 1.5e10; 16: 1.45e10),
 * d = 28 -> 8.4e8 (now 6.2e9; 32: 6.3e9), d = 48 -> 1.5e8 (now 2.5e9), d = 64 -> 4.6e7 (now 1.95e9).  The tensor-core step does not care whether a row of
 * L or U is zero, so the same kernel template runs these dimensions bit-exactly (see the PAD note at the kernel:
 * amh_launch_mh_tc.cu) -- contract v2 only.  A second translation unit so that the two halves compile in parallel. */
#define AMH_MHTC_EXTRA_TU
#include "amh_launch_mh_tc.cu"

namespace amhh {

/* multiples of 8 up to 64, of 16 above (fewer instantiations of the large kernels) */
static int padded_dim(int d) { return d <= 64 ? (d + 7) & ~7 : (d + 15) & ~15; }

bool mh_tc_padded_eligible(const amh_run& r) {
    const amh_sampler& s = *r.sampler;
    const int d = r.dim;
    if (r.cv != AMH_CONTRACT_V2 || r.target->kind != AMH_TARGET_MVNORMAL) return false;
    if (d < 7 || d > 128) return false;
    switch (d) {          /* dimensions with an exact kernel of their own (K1T16 or the per-thread K1) */
    case 8: case 10: case 12: case 16: case 20: case 24: case 32: return false;
    case 9: case 11: case 17: case 18: case 19: return false;      /* per-thread kernel (amh_launch_mh_dims.cu): padding to 16 / 24 costs more */
    }
    if (s.has_mean || s.by_components()) return false;
    if (s.d.cov_kind != AMH_COV_FULL && s.d.cov_kind != AMH_COV_DIAG && s.d.cov_kind != AMH_COV_SCALAR) return false;
    if (s.d.kind == AMH_SAMPLER_STATIC && !s.d.symmetric) return false;      /* needs logq: generic path */
    if (r.pitch % 32 || r.mh_path == 2) return false;
    if (r.x_rows < padded_dim(d)) return false;                              /* the state was not allocated with padding rows */
    static const bool off = std::getenv("AMH_TC_NO_PAD") != nullptr;          /* A/B switch: the generic kernel */
    return !off;
}

/* FEW: only the general-mean variants are instantiated (a zero mean then costs a subtraction of 0 per operand -- the
 * same bits -- instead of a kernel of its own; the kernels above 64 are large) */
/* Below this many chains the 4-warp CTA shape (64 chains per CTA, up to seven CTAs per SM) beats one 28-warp CTA per SM:
 * 448 chains per CTA leave most SMs idle when there are few chains, and a lone CTA is latency-bound at ~6.6 us per step
 * whatever its size.  Measured at d = 32 (profiles/r2_c2_nchains.txt): 1 024 chains 8.8e7 -> 2.5e8, 4 096 3.5e8 -> 1.0e9,
 * 16 384 1.4e9 -> 2.7e9, 32 768 2.8e9 -> 4.2e9 chain-steps/s; at 65 536 the 28-warp CTA wins. */
constexpr long long kSmallRunChains = 49152;
/* (at 65 536 chains, d = 32: one 28-warp CTA per SM 6.34e9 chain-steps/s -- 6.17e9 for this file's padded variant of it --,
 * two 14-warp CTAs 5.49e9, four 7-warp CTAs 5.56e9: profiles/r2_c2_nchains.txt) */

template <int D, int W, bool FEW = false>
static int launch_padded_t(amh_run& r, int nsteps, const SaveArgs& sv) {
    const amh_sampler& s = *r.sampler;
    const amh_target& t = *r.target;
    const int d = r.dim;
    constexpr int NT = (D / 8) * (D / 8 + 1);
    if (!r.scratch) {
        std::vector<double> lf, uf, all;
        if (s.d.cov_kind != AMH_COV_FULL) lf.assign((size_t)NT * 32, 0.0);
        else build_frags(s.scale.data(), d, lf, D);
        build_frags(t.blob.data() + 1 + d, d, uf, D);
        all = lf;
        all.insert(all.end(), uf.begin(), uf.end());
        for (int i = 0; i < D; ++i) all.push_back(i < d ? t.blob[1 + i] : 0.0);                       /* mu, padded */
        for (int i = 0; i < D; ++i)                                                                      /* proposal scales, padded */
            all.push_back(i >= d ? 0.0 : s.d.cov_kind == AMH_COV_DIAG ? s.scale[i] : s.d.cov_kind == AMH_COV_SCALAR ? s.scale[0] : 0.0);
        { const int rca = dmalloc(r.ctx, &r.scratch, all.size() * sizeof(double)); if (rca) return rca; }
        AMH_CUDA_TRY(cudaMemcpyAsync(r.scratch, all.data(), all.size() * sizeof(double), cudaMemcpyHostToDevice, r.ctx->stream));
        AMH_CUDA_TRY(sync_stream(r.ctx, r.ctx->stream));        /* `all` is a stack temporary */
    }
    MhTcArgs a;
    std::memset(&a, 0, sizeof(a));
    a.st = chain_state(r);
    a.sv = sv;
    a.nsteps = nsteps;
    a.is_rw = s.d.kind == AMH_SAMPLER_RW;
    a.mu_zero = FEW ? 0 : 1;
    for (int i = 0; i < d; ++i)
        if (t.blob[1 + i] != 0.0) a.mu_zero = 0;
    a.step0 = (unsigned long long)r.step;
    a.Lf = (const double*)r.scratch;
    a.Uf = a.Lf + (size_t)NT * 32;
    a.mu = a.Uf + (size_t)NT * 32;
    a.c0 = t.blob[0];
    a.dscale = a.mu + D;
    a.d_real = d;
    a.exp_block = (unsigned long long)((d + 3) / 4);            /* contract v2: four normals per block, then the exponential's */
    a.blocks_per_step = a.exp_block + 1ull;
    const bool covd = s.d.cov_kind != AMH_COV_FULL;
    const size_t smem = (size_t)W * tc16_smem_doubles_per_warp<D>() * sizeof(double);
    const unsigned grid = (unsigned)((r.n + 16 * W - 1) / (16 * W));
    const void* key = (const void*)mh_step_tc16_kernel<D, W, false, true, false, 2, true>;
    if (!r.ctx->configured.count(key)) {
#define AMH_TCP_ATTR(...) AMH_CUDA_TRY(cudaFuncSetAttribute(mh_step_tc16_kernel<D, W, __VA_ARGS__, 2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem))
        if constexpr (!FEW) {
            AMH_TCP_ATTR(true, true, false); AMH_TCP_ATTR(true, false, false); AMH_TCP_ATTR(true, true, true); AMH_TCP_ATTR(true, false, true);
        }
        AMH_TCP_ATTR(false, true, false); AMH_TCP_ATTR(false, false, false); AMH_TCP_ATTR(false, true, true); AMH_TCP_ATTR(false, false, true);
#undef AMH_TCP_ATTR
        if constexpr (W <= 7) {
            /* shared memory actually needed by the 28 / W resident CTAs; the rest of the 256 KB stays L1 for the L / U fragments */
            const int need_kb = (int)(((28 / W) * (smem + 1024) + 1023) / 1024);
            const int carve = std::min(100, (need_kb * 100 + 227) / 228 + 1);
#define AMH_TCP_CARVE(...) AMH_CUDA_TRY(cudaFuncSetAttribute(mh_step_tc16_kernel<D, W, __VA_ARGS__, 2, true>, cudaFuncAttributePreferredSharedMemoryCarveout, carve))
            AMH_TCP_CARVE(true, true, false); AMH_TCP_CARVE(true, false, false); AMH_TCP_CARVE(true, true, true); AMH_TCP_CARVE(true, false, true);
            AMH_TCP_CARVE(false, true, false); AMH_TCP_CARVE(false, false, false); AMH_TCP_CARVE(false, true, true); AMH_TCP_CARVE(false, false, true);
#undef AMH_TCP_CARVE
        }
        r.ctx->configured.insert(key);
    }
#define AMH_TCP_GO(...) mh_step_tc16_kernel<D, W, __VA_ARGS__, 2, true><<<grid, 32 * W, smem, r.ctx->stream>>>(a)
    if constexpr (FEW) {
        if (covd) { if (a.is_rw) AMH_TCP_GO(false, true, true); else AMH_TCP_GO(false, false, true); }
        else { if (a.is_rw) AMH_TCP_GO(false, true, false); else AMH_TCP_GO(false, false, false); }
    } else if (covd) {
        if (a.is_rw) { if (a.mu_zero) AMH_TCP_GO(true, true, true); else AMH_TCP_GO(false, true, true); }
        else { if (a.mu_zero) AMH_TCP_GO(true, false, true); else AMH_TCP_GO(false, false, true); }
    } else {
        if (a.is_rw) { if (a.mu_zero) AMH_TCP_GO(true, true, false); else AMH_TCP_GO(false, true, false); }
        else { if (a.mu_zero) AMH_TCP_GO(true, false, false); else AMH_TCP_GO(false, false, false); }
    }
#undef AMH_TCP_GO
    AMH_CUDA_TRY(cudaGetLastError());
    r.launches += 1;
    r.pending_launches += 1;
    return AMH_OK;
}

/* fewer chains than kSmallRunChains, d <= 32 (exact tensor-core dimensions included): the 4-warp CTA shape */
bool mh_tc_small_eligible(const amh_run& r) {
    const amh_sampler& s = *r.sampler;
    const int d = r.dim;
    if (r.cv != AMH_CONTRACT_V2 || r.target->kind != AMH_TARGET_MVNORMAL) return false;
    if (r.n >= kSmallRunChains || d < 7 || d > 32) return false;
    switch (d) {          /* the per-thread kernel K1 has these and spreads its 128-thread CTAs over the SMs already */
    case 9: case 10: case 11: case 12: case 20: return false;      /* (17 ... 19 take the 4-warp tensor-core CTAs when the chains are few) */
    }
    if (s.has_mean || s.by_components()) return false;
    if (s.d.cov_kind != AMH_COV_FULL && s.d.cov_kind != AMH_COV_DIAG && s.d.cov_kind != AMH_COV_SCALAR) return false;
    if (s.d.kind == AMH_SAMPLER_STATIC && !s.d.symmetric) return false;
    if (r.pitch % 32 || r.mh_path == 2 || r.x_rows < padded_dim(d)) return false;
    return std::getenv("AMH_TC_NO_SMALL") == nullptr;                         /* A/B / test switch: always the 28-warp CTA */
}

/* warps per CTA (one CTA per SM) by padded dimension: the per-warp Z / C tile is 160 D + 1 536 bytes */
int launch_mh_tc_padded(amh_run& r, int nsteps, const SaveArgs& sv) {
    if (mh_tc_small_eligible(r)) {
        switch (padded_dim(r.dim)) {
        case 8: return launch_padded_t<8, 4>(r, nsteps, sv);
        case 16: return launch_padded_t<16, 4>(r, nsteps, sv);
        case 24: return launch_padded_t<24, 4>(r, nsteps, sv);
        case 32: return launch_padded_t<32, 4>(r, nsteps, sv);
        }
    }
    /* D = 40 ... 64 with so few chains that the large CTAs would cover less than half of the SMs: 8-warp CTAs (128 chains) */
    if (r.dim > 32 && r.dim <= 64 && !std::getenv("AMH_TC_NO_SMALL")) {
        const int D = padded_dim(r.dim);
        const int wbig = D == 40 ? 24 : D == 48 ? 20 : 16;
        /* measured (profiles/r2_c2_nchains.txt): D = 40 is faster on 8-warp CTAs at every size tried (65 536 chains: 3.6e9 vs
         * 2.7e9 -- 171 CTAs of 384 chains are 1.15 waves), D = 48 up to ~100 000 chains (2.75e9 vs 2.5e9), D = 56 / 64 only
         * while the large CTAs would leave SMs idle */
        long long small_max = D == 40 ? (1ll << 62) : D == 48 ? 100000 : 74ll * 16 * wbig;
        if (const char* ev = std::getenv("AMH_TC_SMALL_MAX")) small_max = std::atoll(ev);      /* A/B switch */
        if (r.n < small_max) {
            switch (D) {
            case 40: return launch_padded_t<40, 8, true>(r, nsteps, sv);
            case 48: return launch_padded_t<48, 8, true>(r, nsteps, sv);
            case 56: return launch_padded_t<56, 8, true>(r, nsteps, sv);
            case 64: return launch_padded_t<64, 8, true>(r, nsteps, sv);
            }
        }
    }
    switch (padded_dim(r.dim)) {
    case 8: return launch_padded_t<8, 28>(r, nsteps, sv);
    case 16: return launch_padded_t<16, 28>(r, nsteps, sv);
    case 24: return launch_padded_t<24, 28>(r, nsteps, sv);
    case 32: return launch_padded_t<32, 28>(r, nsteps, sv);
    case 40: return launch_padded_t<40, 24>(r, nsteps, sv);
    case 48: return launch_padded_t<48, 20>(r, nsteps, sv);
    case 56: return launch_padded_t<56, 16>(r, nsteps, sv);
    case 64: return launch_padded_t<64, 16>(r, nsteps, sv);
    case 80: return launch_padded_t<80, 14, true>(r, nsteps, sv);
    case 96: return launch_padded_t<96, 12, true>(r, nsteps, sv);
    case 112: return launch_padded_t<112, 10, true>(r, nsteps, sv);
    case 128: return launch_padded_t<128, 10, true>(r, nsteps, sv);
    }
    return fail(AMH_ERR_INVALID, "padded tensor-core MH path: unsupported dimension");
}

}  // namespace amhh
