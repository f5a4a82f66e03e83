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
#include <initializer_list>
#include <type_traits>
#include "amh_params.cuh"

namespace amhh {
using namespace amhd;

/* @rtc-begin: the device code from here to @rtc-end is also compiled by NVRTC for user-supplied targets (amh_rtc.cu) */
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
                    save_moments(a.sv, o, v);
                }
            }
            if (a.sv.out) a.sv.out[(long long)d * a.sv.out_pitch + ch] = lpold[ch];
            if (a.sv.acc_out) a.sv.acc_out[ch] = a.st.acc[ch];
        }
    }
}

/* ---------------------------------------------------------------------------
 * K2F: same exact sequential semantics, scheduled ahead of time and executed as a dataflow graph.
 *
 * Everything of a sweep that does not depend on the walker positions is known before the sweep runs: the partner
 * index, the stretch factor z, (d-1) log z, the exponential -- and therefore also the whole DEPENDENCY FOREST
 * ("walker i needs the new value of partner idx_i < i").  Two kernels per launch:
 *
 *   stretch_plan_kernel   one CTA per (sweep, ensemble), on all SMs: draws the partners, computes every walker's
 *                         LEVEL in the forest (0 if idx > i, else level(idx) + 1; a pointer chase through shared
 *                         memory, expected length e - 1), counting-sorts the walkers by level -- every level padded
 *                         to whole warps with sentinel slots -- and stores, in slot order, (self, partner) packed
 *                         in 32 bits, z, (d-1) log z and the exponential.
 *   stretch_sweep_flow_kernel   one CTA per ensemble.  Warps take 32-slot chunks in slot order; a lane whose partner
 *                         is a NEW value spins on that walker's version flag in shared memory (`ver[j]` = last sweep
 *                         walker j finished), executes its move, fences and publishes its own flag.  A chunk never
 *                         mixes levels, so every dependency points to a lower slot owned by a warp that reaches it
 *                         first: the lowest unfinished chunk can always run -- no deadlock, no barrier per level,
 *                         one __syncthreads per sweep (it protects the old buffer from the next sweep's writes).
 *                         Walker data stays in L2 (ld/st .cg); plan entries are prefetched one chunk ahead.  NEW values
 *                         that a later walker of the same sweep reads (e^-1 of the walkers; the plan knows which)
 *                         are also FORWARDED through shared memory, so a level of the dependency chain costs the
 *                         move itself plus a shared-memory hop instead of a store->load round trip through L2.
 *
 * Levels >= lcap - 1 (lcap = kStretchLevels; never reached in practice: a chain of length L has probability ~ 1/L!)
 * share the last bucket, which one thread executes in increasing walker order -- the sequential sweep itself. */
constexpr int kStretchLevels = 32;
constexpr unsigned kStretchSentinel = 0xffffffffu;

struct StretchPlan {
    unsigned* pair;            /* [nsteps][n_ensembles][nwp] slot order: self | partner << 16, or the sentinel */
    unsigned* fwd;             /* same indexing: forwarding slot of self (0xffff: nobody reads the new value in this
                                  sweep, 0xfffe: somebody does, through L2) | of the partner << 16 (0xffff = L2) */
    double* zf;                /* same indexing */
    double* am;
    double* ex;
    int* meta;                 /* [nsteps][n_ensembles][4]: padded slots of the parallel levels, overflow lo, hi, - */
    long long nwp;             /* slots per (sweep, ensemble) = roundup(nw, 32) + 32 * kStretchLevels */
};

template <int BLOCK>
__global__ void __launch_bounds__(BLOCK)
stretch_plan_kernel(StretchPlan o, const unsigned long long* __restrict__ seeds, long long n, int nw, int d,
                    unsigned long long step0, double aa, int lcap, int fcap) {
    extern __shared__ int smem_pl[];
    int* partner = smem_pl;                    /* [nw] */
    int* slotinfo = partner + nw;              /* [nw] bucket << 16 | rank within the bucket */
    int* fslot = slotinfo + nw;                /* [nw] 1 = some later walker reads this walker's new value -> slot */
    __shared__ int hist[kStretchLevels];
    __shared__ int start[kStretchLevels + 1];
    __shared__ int wsum[BLOCK / 32];
    const int tid = threadIdx.x;
    const long long nens = n / nw;
    const int s = (int)(blockIdx.x / nens);
    const long long en = blockIdx.x % nens;
    const unsigned long long k = step0 + (unsigned long long)s + 1ull;
    const unsigned long long seed = seeds[en];
    if (tid < kStretchLevels) hist[tid] = 0;
    for (int i = tid; i < nw; i += BLOCK) {
        const unsigned long long blk = (k * (unsigned long long)nw + (unsigned long long)i) * 2ull;
        const amh::Block b0 = amh::stream_block(seed, blk, 0u);
        /* idx = mod1(i + rand(1:(n-1)), n)  (emcee.jl:52) */
        const long long rr = (long long)amh::bounded(b0.v[0], b0.v[1], (unsigned long long)(nw - 1));
        partner[i] = (int)((i + rr + 1) % nw);
        fslot[i] = 0;
    }
    __syncthreads();
    for (int i = tid; i < nw; i += BLOCK) {
        int lvl = 0, cur = i, j = partner[i];
        if (j < i) fslot[j] = 1;               /* benign race: everybody writes 1 */
        while (j < cur && lvl < lcap - 1) { ++lvl; cur = j; j = partner[cur]; }
        slotinfo[i] = (lvl << 16) | atomicAdd(&hist[lvl], 1);
    }
    __syncthreads();
    if (tid == 0) {
        int acc = 0;
        for (int l = 0; l < kStretchLevels; ++l) { start[l] = acc; acc += (hist[l] + 31) & ~31; }
        start[kStretchLevels] = acc;
    }
    /* forwarding slots: exclusive scan of the flags (thread t owns the contiguous walkers [t*per, (t+1)*per)) */
    {
        const int per = (nw + BLOCK - 1) / BLOCK;
        const int lo = tid * per, hi = min(nw, lo + per);
        int cnt = 0;
        for (int i = lo; i < hi; ++i) cnt += fslot[i];
        int inc = cnt;
        for (int o2 = 1; o2 < 32; o2 <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, inc, o2);
            if ((tid & 31) >= o2) inc += v;
        }
        if ((tid & 31) == 31) wsum[tid >> 5] = inc;
        __syncthreads();
        if (tid < 32) {
            int v = (tid < BLOCK / 32) ? wsum[tid] : 0;
            for (int o2 = 1; o2 < 32; o2 <<= 1) {
                const int u = __shfl_up_sync(0xffffffffu, v, o2);
                if (tid >= o2) v += u;
            }
            if (tid < BLOCK / 32) wsum[tid] = v;           /* inclusive over warps */
        }
        __syncthreads();
        int run = inc - cnt + ((tid >> 5) ? wsum[(tid >> 5) - 1] : 0);
        for (int i = lo; i < hi; ++i) {
            const int f = fslot[i];
            fslot[i] = f ? (run < fcap ? run : 0xfffe) : 0xffff;     /* 0xfffe: read later in this sweep, through L2 */
            run += f;
        }
    }
    __syncthreads();
    const size_t off = ((size_t)s * nens + en) * (size_t)o.nwp;
    if (tid == 0) {
        int* m = o.meta + ((size_t)s * nens + en) * 4;
        m[0] = start[lcap - 1];                               /* padded slots of the parallel levels */
        m[1] = start[lcap - 1];                               /* overflow bucket [lo, hi) */
        m[2] = start[lcap - 1] + hist[lcap - 1];
        m[3] = min(wsum[BLOCK / 32 - 1], fcap);               /* forwarding slots in use */
    }
    /* sentinels in the padding of every level */
    for (int l = tid >> 5; l < kStretchLevels; l += BLOCK >> 5) {
        const int q = start[l] + hist[l] + (tid & 31);
        if (q < start[l + 1]) { o.pair[off + q] = kStretchSentinel; o.fwd[off + q] = 0xffffffffu; }
    }
    for (int i = tid; i < nw; i += BLOCK) {
        const unsigned long long blk = (k * (unsigned long long)nw + (unsigned long long)i) * 2ull;
        const amh::Block b0 = amh::stream_block(seed, blk, 0u);
        const amh::Block b1 = amh::stream_block(seed, blk + 1ull, 0u);
        const int si = slotinfo[i];
        const size_t slot = off + start[si >> 16] + (si & 0xffff);
        const double u = amh::u01(b0.v[2], b0.v[3]);
        const double tt = (aa - 1.0) * u + 1.0;
        const double z = (tt * tt) / aa;
        const int pj = partner[i];
        o.pair[slot] = (unsigned)i | ((unsigned)pj << 16);
        o.fwd[slot] = (unsigned)fslot[i] | (((pj < i && fslot[pj] < 0xfffe) ? (unsigned)fslot[pj] : 0xffffu) << 16);
        o.zf[slot] = z;
        o.am[slot] = (double)(d - 1) * amh::log_(z);
        o.ex[slot] = amh::exponential(b1.v[0], b1.v[1]);
    }
}

/* 256-bit global accesses through L2 (sm_100: LDG/STG.E.ENL2.256) */
__device__ __forceinline__ void ld256(const double* p, double& a, double& b, double& c, double& d) {
#if defined(AMH_NO_V4_F64)      /* run-time compilation with an NVRTC older than CUDA 12.9 (amh_rtc.cu) */
    asm volatile("ld.global.cg.v2.f64 {%0,%1}, [%2];" : "=d"(a), "=d"(b) : "l"(p) : "memory");
    asm volatile("ld.global.cg.v2.f64 {%0,%1}, [%2];" : "=d"(c), "=d"(d) : "l"(p + 2) : "memory");
#else
    asm volatile("ld.global.cg.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(a), "=d"(b), "=d"(c), "=d"(d) : "l"(p) : "memory");
#endif
}
__device__ __forceinline__ void st256(double* p, double a, double b, double c, double d) {
#if defined(AMH_NO_V4_F64)
    asm volatile("st.global.cg.v2.f64 [%2], {%0,%1};" :: "d"(a), "d"(b), "l"(p) : "memory");
    asm volatile("st.global.cg.v2.f64 [%2], {%0,%1};" :: "d"(c), "d"(d), "l"(p + 2) : "memory");
#else
    asm volatile("st.global.cg.v4.f64 [%4], {%0,%1,%2,%3};" :: "d"(a), "d"(b), "d"(c), "d"(d), "l"(p) : "memory");
#endif
}

/* Inside a launch the walkers of an ensemble live as RECORDS [x_0 .. x_{d-1}, lp, pad] of RS = roundup(d + 1, 4)
 * doubles (32-byte multiples), double-buffered (old sweep / new sweep): a move touches 3 records with 256-bit
 * accesses -- RS/4 128-byte-line requests per lane and record instead of one per coordinate in the [dim][chain]
 * layout, and it is the L1 line-request rate of the ensemble's SM that bounds the sweep (random partners). */
template <int DMAX>
struct Rec {
    static constexpr int cap = Dim<DMAX>::fixed ? ((DMAX + 1 + 3) & ~3) : ((kGenericCap + 1 + 3) & ~3);
    __host__ __device__ static int size(int d) { return (d + 1 + 3) & ~3; }
};

/* shared memory of the sweep kernel */
struct StretchSmem {
    double* fwd;                          /* [fcap][d] forwarded new positions */
    unsigned* naccs;                      /* [nw] accepted moves of this launch */
    volatile unsigned short* ver;         /* [nw] last sweep (1-based, this launch) the walker finished */
    unsigned char* accs;                  /* [nw] last accept flag */
    /* 2-CTA cluster variant (one ensemble on two SMs): every CTA keeps its own copy of `fwd` and `ver`, the mover of a
     * walker writes both (the peer's through distributed shared memory); naccs / accs live in rank 0 only */
    volatile unsigned short* ver_peer;
    /* forwarded values cross the cluster WITHOUT a cluster-scope fence (fence.acq_rel.cluster = MEMBAR.ALL.GPU + an L1
     * invalidation on the acquire side: ~3 000 cycles on the critical path of every level of the dependency forest):
     * every forwarding slot has one mbarrier per CTA that completes exactly once per sweep.  The mover of the walker
     * completes its OWN CTA's with a plain arrive (release.cta orders its st.shared of the value) and the PEER's with a
     * remote arrive.expect_tx followed by st.async of the value -- the async stores carry complete_tx, so the peer's
     * barrier completes when, and only when, all bytes have landed in the peer's copy of the slot.  Readers only ever
     * wait on a barrier of their own CTA.  Shared-memory addresses below are 32-bit shared::cta / shared::cluster. */
    unsigned bar_l;                       /* this CTA's barriers [fmax] */
    unsigned bar_r;                       /* the peer's */
    unsigned fwd_r;                       /* the peer's copy of `fwd` */
};

__device__ __forceinline__ unsigned cluster_ctarank() {
    unsigned r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void fence_cluster() { asm volatile("fence.acq_rel.cluster;" ::: "memory"); }
/* acquire side of a flag that a thread of the peer CTA raised after a cluster-scope release fence: unlike a fence it
 * does not wait for this thread's own outstanding stores */
__device__ __forceinline__ unsigned short ld_acquire_cluster_u16(const volatile unsigned short* p) {
    unsigned short v;
    asm volatile("ld.acquire.cluster.shared::cta.u16 %0, [%1];" : "=h"(v) : "r"((unsigned)__cvta_generic_to_shared((const void*)p)) : "memory");
    return v;
}
/* generic address of the same shared-memory location in CTA `rank` of the cluster */
template <class P>
__device__ __forceinline__ P* map_cta(P* p, unsigned rank) {
    unsigned long long in = (unsigned long long)p, out;
    asm volatile("mapa.u64 %0, %1, %2;" : "=l"(out) : "l"(in), "r"(rank));
    return (P*)out;
}

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ unsigned map_cta_u32(unsigned addr, unsigned rank) {
    unsigned out;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(out) : "r"(addr), "r"(rank));
    return out;
}
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned bar) {                 /* release.cta: orders this thread's st.shared */
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx_peer(unsigned bar_cluster, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cluster.b64 _, [%0], %1;" :: "r"(bar_cluster), "r"(bytes) : "memory");
}
__device__ __forceinline__ void st_async_peer(unsigned dst_cluster, double a, unsigned bar_cluster) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.f64 [%0], %1, [%2];"
                 :: "r"(dst_cluster), "d"(a), "r"(bar_cluster) : "memory");
}
__device__ __forceinline__ void st_async_peer2(unsigned dst_cluster, double a, double b, unsigned bar_cluster) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.f64 [%0], {%1,%2}, [%3];"
                 :: "r"(dst_cluster), "d"(a), "d"(b), "r"(bar_cluster) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {  /* acquire.cta; try_wait sleeps in hardware */
    unsigned ok;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    } while (!ok);
}

/* one stretch move (emcee.jl:70-102) of walker `i` with partner `idx`.  `w` = the walker's own record (already loaded);
 * the partner comes from the forwarding slot `pslot` (shared memory) if it has one, else from its record in L2;
 * the result goes to the walker's new record and, if somebody reads it later in this sweep, to slot `sslot`. */
template <int DMAX, class T, int CL = 1>
__device__ __forceinline__ void stretch_move(const typename T::template Params<DMAX>& tp, int d, long long base, int i,
                                             int idx, unsigned sslot, unsigned pslot, double z, double am, double ex,
                                             double (&w)[Rec<DMAX>::cap], const double* __restrict__ Rold,
                                             double* __restrict__ Rnew, const StretchSmem& sm, unsigned short want) {
    using D = Dim<DMAX>;
    constexpr int CAP = D::cap;
    constexpr int RC = Rec<DMAX>::cap;
    const int rs = D::fixed ? RC : Rec<DMAX>::size(d);
    double o[RC], y[CAP];
    if (pslot != 0xffffu) {
        const double* ps = sm.fwd + (size_t)pslot * d;
        if constexpr (D::fixed) {
#pragma unroll
            for (int j = 0; j < DMAX; ++j) o[j] = ps[j];
        } else {
            for (int j = 0; j < d; ++j) o[j] = ps[j];
        }
    } else {
        const double* po = ((idx < i) ? Rnew : Rold) + (size_t)(base + idx) * rs;      /* emcee.jl:53 */
        if constexpr (D::fixed) {
#pragma unroll
            for (int j = 0; j < ((DMAX + 3) & ~3); j += 4) ld256(po + j, o[j], o[j + 1], o[j + 2], o[j + 3]);
        } else {
            for (int j = 0; j < d; j += 4) ld256(po + j, o[j], o[j + 1], o[j + 2], o[j + 3]);
        }
    }
    if constexpr (D::fixed) {
#pragma unroll
        for (int j = 0; j < DMAX; ++j) y[j] = o[j] + z * (w[j] - o[j]);
    } else {
        for (int j = 0; j < d; ++j) y[j] = o[j] + z * (w[j] - o[j]);
    }
    const double lpw = w[d];
    const double lpy = T::template logp<DMAX>(y, d, tp);
    const double alpha = (am + lpy) - lpw;
    const bool acc = (-ex <= alpha);                         /* emcee.jl:93 (non-strict) */
    double* pn = Rnew + (size_t)(base + i) * rs;
    if constexpr (D::fixed) {
#pragma unroll
        for (int j = 0; j < DMAX; ++j) w[j] = acc ? y[j] : w[j];
        w[DMAX] = acc ? lpy : lpw;
    } else {
        for (int j = 0; j < d; ++j) w[j] = acc ? y[j] : w[j];
        w[d] = acc ? lpy : lpw;
    }
    /* publish: a forwarded value only has to be ordered against the shared-memory copy, so the flag goes up before
     * the record's round trip to L2 (which the next sweep needs, and the sweep's closing barrier orders) */
    if (sslot < 0xfffeu) {
        double* ps = sm.fwd + (size_t)sslot * d;
        if constexpr (D::fixed) {
#pragma unroll
            for (int j = 0; j < DMAX; ++j) ps[j] = w[j];
        } else {
            for (int j = 0; j < d; ++j) ps[j] = w[j];
        }
        if constexpr (CL == 2) {
            /* the peer's copy first (the longer trip): expect_tx + async stores that complete the peer's barrier */
            const unsigned rb = sm.bar_r + sslot * 8u;
            const unsigned rf = sm.fwd_r + sslot * (unsigned)d * 8u;
            mbar_arrive_expect_tx_peer(rb, (unsigned)d * 8u);
            if constexpr (D::fixed && (DMAX % 2 == 0)) {
#pragma unroll
                for (int j = 0; j < DMAX; j += 2) st_async_peer2(rf + j * 8u, w[j], w[j + 1], rb);
            } else if constexpr (D::fixed) {
#pragma unroll
                for (int j = 0; j < DMAX; ++j) st_async_peer(rf + j * 8u, w[j], rb);
            } else {
                for (int j = 0; j < d; ++j) st_async_peer(rf + j * 8u, w[j], rb);
            }
            mbar_arrive(sm.bar_l + sslot * 8u);              /* own CTA's barrier: after the st.shared above */
        } else {
            __threadfence_block();
            sm.ver[i] = want;
        }
    }
    if constexpr (D::fixed) {
#pragma unroll
        for (int j = 0; j < RC; j += 4) st256(pn + j, w[j], w[j + 1], w[j + 2], w[j + 3]);
    } else {
        for (int j = 0; j < rs; j += 4) st256(pn + j, w[j], w[j + 1], w[j + 2], w[j + 3]);
    }
    if (sslot == 0xfffeu) {                                  /* read through L2 later in this sweep */
        if constexpr (CL == 2) {
            fence_cluster();
            sm.ver[i] = want;
            sm.ver_peer[i] = want;
        } else {
            __threadfence_block();
            sm.ver[i] = want;
        }
    }
    sm.accs[i] = acc ? 1 : 0;
    if constexpr (CL == 2) {
        if (acc) atomicAdd(sm.naccs + i, 1u);                /* rank 0's counter, possibly remote: fire and forget */
    } else {
        if (acc) sm.naccs[i] += 1u;                          /* only ever touched by the thread that moves walker i */
    }
}

template <int DMAX>
__device__ __forceinline__ void stretch_load_own(double (&w)[Rec<DMAX>::cap], const double* __restrict__ Rold, long long base,
                                                 int i, int d) {
    constexpr int RC = Rec<DMAX>::cap;
    if constexpr (Dim<DMAX>::fixed) {
        const double* pw = Rold + (size_t)(base + i) * RC;
#pragma unroll
        for (int j = 0; j < RC; j += 4) ld256(pw + j, w[j], w[j + 1], w[j + 2], w[j + 3]);
    } else {
        const int rs = Rec<DMAX>::size(d);
        const double* pw = Rold + (size_t)(base + i) * rs;
        for (int j = 0; j < rs; j += 4) ld256(pw + j, w[j], w[j + 1], w[j + 2], w[j + 3]);
    }
}

template <int DMAX, class T, int BLOCK, int CL>
__device__ __forceinline__ void stretch_flow_body(const StretchArgs& a, const StretchPlan& plan,
                                                  const typename T::template Params<DMAX>& tp, double* RA, double* RB, int fcap) {
    using D = Dim<DMAX>;
    extern __shared__ __align__(16) double smem_fl[];
    const int nw = (int)a.n_walkers;
    const int d = D::fixed ? DMAX : a.d;
    StretchSmem sm;
    sm.fwd = smem_fl;
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(sm.fwd + (size_t)fcap * d);    /* [fcap], cluster variant only */
    sm.naccs = reinterpret_cast<unsigned*>(bars + (CL == 2 ? fcap : 0));
    unsigned short* ver_nv = reinterpret_cast<unsigned short*>(sm.naccs + nw);
    sm.ver = ver_nv;
    sm.accs = reinterpret_cast<unsigned char*>(ver_nv + nw);
    sm.ver_peer = nullptr;
    sm.bar_l = sm.bar_r = sm.fwd_r = 0u;
    const unsigned rank = (CL == 2) ? cluster_ctarank() : 0u;
    if constexpr (CL == 2) {
        sm.ver_peer = map_cta(ver_nv, rank ^ 1u);
        sm.bar_l = smem_u32(bars);
        sm.bar_r = map_cta_u32(sm.bar_l, rank ^ 1u);
        sm.fwd_r = map_cta_u32(smem_u32(sm.fwd), rank ^ 1u);
        if (rank != 0u) {
            sm.naccs = map_cta(sm.naccs, 0u);
            sm.accs = map_cta(sm.accs, 0u);
        }
    }
    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    constexpr int NWARP = BLOCK / 32;
    const long long en = blockIdx.x / CL;
    const long long nens = a.st.n / nw;
    const long long base = en * nw;
    const int rs = D::fixed ? Rec<DMAX>::cap : Rec<DMAX>::size(d);
    double* Rold = RA;  double* Rnew = RB;
    /* prologue: [dim][chain] state -> records */
    int fmax = 0;                                                         /* forwarding slots any sweep of this launch uses */
    if constexpr (CL == 2) {
        for (int i = tid; i < nw; i += BLOCK) ver_nv[i] = 0;              /* every CTA clears its own copy */
        for (int s = 0; s < a.nsteps; ++s) fmax = max(fmax, plan.meta[((size_t)s * nens + en) * 4 + 3]);
        for (int q = tid; q < fmax; q += BLOCK) mbar_init(sm.bar_l + q * 8u, 1u);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = tid + (int)rank * BLOCK; i < nw; i += CL * BLOCK) {
        ver_nv[i] = 0;
        sm.naccs[i] = 0u;
        sm.accs[i] = a.st.acc[base + i];
        double* rec = RA + (size_t)(base + i) * rs;
        for (int j = 0; j < d; ++j) rec[j] = a.st.X[(long long)j * a.st.pitch + base + i];
        rec[d] = a.st.lp[base + i];
        for (int j = d + 1; j < rs; ++j) rec[j] = 0.0;
    }
    const int* __restrict__ meta0 = plan.meta + (size_t)en * 4;
    int npar = a.nsteps > 0 ? meta0[0] : 0, olo = a.nsteps > 0 ? meta0[1] : 0, ohi = a.nsteps > 0 ? meta0[2] : 0;
    int nfwd = a.nsteps > 0 ? meta0[3] : 0;
    if constexpr (CL == 2) { __threadfence(); cluster_sync_all(); } else __syncthreads();

    for (int s = 0; s < a.nsteps; ++s) {
        const size_t off = ((size_t)s * nens + en) * (size_t)plan.nwp;
        const unsigned short want = (unsigned short)(s + 1);
        /* the next sweep's meta data, long before it is needed */
        int npar_n = 0, olo_n = 0, ohi_n = 0, nfwd_n = 0;
        if (s + 1 < a.nsteps) {
            const int* __restrict__ mn = plan.meta + ((size_t)(s + 1) * nens + en) * 4;
            npar_n = mn[0]; olo_n = mn[1]; ohi_n = mn[2]; nfwd_n = mn[3];
        }
        const unsigned par = (unsigned)s & 1u;                          /* every slot barrier completes once per sweep */
        /* plan entries one chunk ahead: they do not depend on the walkers */
        int q = ((int)rank * NWARP + warp) * 32 + lane;
        unsigned pr = kStretchSentinel, fw = 0xffffffffu;
        double z = 0.0, am = 0.0, ex = 0.0;
        if (q < npar) {
            pr = __ldg(plan.pair + off + q); fw = __ldg(plan.fwd + off + q); z = __ldg(plan.zf + off + q);
            am = __ldg(plan.am + off + q);   ex = __ldg(plan.ex + off + q);
        }
#pragma unroll 1
        for (; q - lane < npar; q += CL * NWARP * 32) {
            const unsigned pr_c = pr, fw_c = fw;
            const double z_c = z, am_c = am, ex_c = ex;
            const int qn = q + CL * NWARP * 32;
            if (qn < npar) {
                pr = __ldg(plan.pair + off + qn); fw = __ldg(plan.fwd + off + qn); z = __ldg(plan.zf + off + qn);
                am = __ldg(plan.am + off + qn);   ex = __ldg(plan.ex + off + qn);
            } else {
                pr = kStretchSentinel;
            }
            if (pr_c != kStretchSentinel) {
                const int i = (int)(pr_c & 0xffffu), idx = (int)(pr_c >> 16);
                double w[Rec<DMAX>::cap];
                stretch_load_own<DMAX>(w, Rold, base, i, d);             /* in flight while the lane waits */
                if (idx < i) {                                           /* the partner's new value (emcee.jl:53) */
                    if (CL == 2 && (fw_c >> 16) != 0xffffu) {
                        mbar_wait(sm.bar_l + (fw_c >> 16) * 8u, par);        /* forwarded: this CTA's barrier of the slot */
                    } else {
                        while (sm.ver[idx] != want) { }
                        if constexpr (CL == 2) (void)ld_acquire_cluster_u16(sm.ver + idx); else __threadfence_block();
                    }
                }
                stretch_move<DMAX, T, CL>(tp, d, base, i, idx, fw_c & 0xffffu, fw_c >> 16, z_c, am_c, ex_c, w, Rold, Rnew, sm, want);
            }
        }
        /* cluster variant: before anybody may start the next sweep every slot barrier of this CTA must have completed
         * its phase -- used slots by their mover (wait: the peer's async stores may still be in flight), unused ones here */
        auto close_slots = [&]() {
            if constexpr (CL == 2) {
                for (int q = tid; q < fmax; q += BLOCK) {
                    if (q >= nfwd) mbar_arrive(sm.bar_l + q * 8u);       /* unused in this sweep: completes the phase at once */
                    mbar_wait(sm.bar_l + q * 8u, par);
                }
            }
        };
        if (ohi <= olo) close_slots();
        if constexpr (CL == 2) { __threadfence(); cluster_sync_all(); } else __syncthreads();
        /* overflow bucket: in increasing walker order by one thread = the reference's own loop */
        if (ohi > olo) {
            if (tid == 0 && rank == 0u) {
                int last = -1;
                for (int c = olo; c < ohi; ++c) {
                    int best = 0x7fffffff, bq = olo;
                    for (int qq = olo; qq < ohi; ++qq) {
                        const int self = (int)(plan.pair[off + qq] & 0xffffu);
                        if (self > last && self < best) { best = self; bq = qq; }
                    }
                    const unsigned po = plan.pair[off + bq], fo = plan.fwd[off + bq];
                    double w[Rec<DMAX>::cap];
                    stretch_load_own<DMAX>(w, Rold, base, (int)(po & 0xffffu), d);
                    if (CL == 2 && (int)(po >> 16) < (int)(po & 0xffffu) && (fo >> 16) != 0xffffu)
                        mbar_wait(sm.bar_l + (fo >> 16) * 8u, par);     /* a value the peer forwarded may still be landing */
                    stretch_move<DMAX, T, CL>(tp, d, base, (int)(po & 0xffffu), (int)(po >> 16), fo & 0xffffu, fo >> 16,
                                              plan.zf[off + bq], plan.am[off + bq], plan.ex[off + bq], w, Rold, Rnew, sm, want);
                    __threadfence_block();
                    last = best;
                }
            }
            if constexpr (CL == 2) {
                __syncthreads();                                        /* the bucket's own forwarding is done */
                close_slots();
                __threadfence(); cluster_sync_all();
            } else __syncthreads();
        }
        double* tR = Rold; Rold = Rnew; Rnew = tR;
        npar = npar_n; olo = olo_n; ohi = ohi_n; nfwd = nfwd_n;
    }
    /* epilogue: records -> [dim][chain] state (always the run's primary buffers), counters, save point outputs */
    for (int i = tid + (int)rank * BLOCK; i < nw; i += CL * BLOCK) {
        const long long ch = base + i;
        const double* rec = Rold + (size_t)ch * rs;
        for (int j = 0; j < d; ++j) {
            const double v = __ldcg(rec + j);
            const long long o = (long long)j * a.st.pitch + ch;
            a.st.X[o] = v;
            if (a.sv.out) a.sv.out[(long long)j * a.sv.out_pitch + ch] = v;
            if (a.sv.sum) {
                save_moments(a.sv, o, v);
            }
        }
        const double lpv = __ldcg(rec + d);
        a.st.lp[ch] = lpv;
        a.st.acc[ch] = sm.accs[i];
        a.st.nacc[ch] = a.st.nacc[ch] + (unsigned long long)sm.naccs[i];
        if (a.sv.out) a.sv.out[(long long)d * a.sv.out_pitch + ch] = lpv;
        if (a.sv.acc_out) a.sv.acc_out[ch] = sm.accs[i];
    }
    if constexpr (CL == 2) cluster_sync_all();             /* rank 0's shared memory must outlive the peer's reads */
}

template <int DMAX, class T, int BLOCK>
__global__ void __launch_bounds__(BLOCK, 1)
stretch_sweep_flow_kernel(const __grid_constant__ StretchArgs a, const __grid_constant__ StretchPlan plan,
                          const __grid_constant__ typename T::template Params<DMAX> tp, double* RA, double* RB, int fcap) {
    stretch_flow_body<DMAX, T, BLOCK, 1>(a, plan, tp, RA, RB, fcap);
}

/* the same sweep with one ensemble on a CLUSTER OF TWO CTAs (two SMs, two L1s): chunks alternate between the 2 x NWARP
 * warps of the cluster in slot order, so every dependency still points to a chunk that a resident warp owns; flags and
 * forwarded values are mirrored into the peer's shared memory (DSMEM), sweeps end in a cluster barrier.  Used when two
 * CTAs per ensemble still fit the GPU in one wave (BASELINE config 3: 64 ensembles -> 128 of 148 SMs). */
template <int DMAX, class T, int BLOCK>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(BLOCK, 1)
stretch_sweep_flow2_kernel(const __grid_constant__ StretchArgs a, const __grid_constant__ StretchPlan plan,
                           const __grid_constant__ typename T::template Params<DMAX> tp, double* RA, double* RB, int fcap) {
    stretch_flow_body<DMAX, T, BLOCK, 2>(a, plan, tp, RA, RB, fcap);
}

/* @rtc-end */
}  // namespace amhh
#include "amh_fastmath.cuh"
#include "amh_launch_stretch_res.cuh"
namespace amhh {

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
    if constexpr (DMAX > 0 && T::kind != AMH_TARGET_USER) {          /* K2R: the ensemble resident in a 2-CTA cluster */
        int rc = AMH_OK;
        if (launch_stretch_res_t<DMAX, T>(r, nsteps, a, tp, &rc)) return rc;
    }
    if (a.n_walkers <= 16384 && nsteps < 65535) {   /* 16-bit walker indices / sweep versions, 8 B of shared memory per walker */
        const long long nens = r.n / a.n_walkers;
        StretchPlan plan;
        plan.nwp = ((a.n_walkers + 31) & ~31ll) + 32 * kStretchLevels;
        const size_t slots = (size_t)nsteps * nens * plan.nwp;
        const size_t recs = (size_t)r.n * (size_t)Rec<DMAX>::size(r.dim);          /* doubles per record buffer */
        /* one plan buffer: zf, am, ex (doubles), pair, fwd (unsigned), meta (4 ints per sweep and ensemble), 256-byte aligned */
        const size_t pbytes = (slots * (3 * sizeof(double) + 2 * sizeof(unsigned)) + (size_t)nsteps * nens * 4 * sizeof(int) + 255) & ~(size_t)255;
        const size_t need = 2 * recs * sizeof(double) + 256 + 2 * pbytes;
        if (!r.aux_stream) {
            AMH_CUDA_TRY(cudaStreamCreateWithFlags(&r.aux_stream, cudaStreamNonBlocking));
            for (int b = 0; b < 2; ++b) {
                AMH_CUDA_TRY(cudaEventCreateWithFlags(&r.ev_plan[b], cudaEventDisableTiming));
                AMH_CUDA_TRY(cudaEventCreateWithFlags(&r.ev_sweep[b], cudaEventDisableTiming));
            }
        }
        if (r.plan_layout_nsteps != nsteps) {
            /* the two plan buffers are laid out for one launch length: on a change nothing in flight may touch them */
            AMH_CUDA_TRY(cudaStreamSynchronize(r.ctx->stream));
            AMH_CUDA_TRY(cudaStreamSynchronize(r.aux_stream));
            r.plan_step0[0] = r.plan_step0[1] = -1;
            r.plan_layout_nsteps = nsteps;
        }
        if (need > r.scratch_bytes) {
            AMH_CUDA_TRY(cudaStreamSynchronize(r.aux_stream));      /* a plan made ahead may still be running */
            r.plan_step0[0] = r.plan_step0[1] = -1;
            dfree(r.ctx, r.scratch);
            r.scratch = nullptr; r.scratch_bytes = 0;
            const int rca = dmalloc(r.ctx, &r.scratch, need);
            if (rca) return rca;
            r.scratch_bytes = need;
        }
        double* RA = (double*)r.scratch;                 /* record buffers first: 256-byte aligned pool memory */
        double* RB = RA + recs;
        char* pbase = (char*)(((uintptr_t)(RB + recs) + 255) & ~(uintptr_t)255);
        auto plan_at = [&](int b) {
            StretchPlan q = plan;
            char* p0 = pbase + (size_t)b * pbytes;
            q.zf = (double*)p0;
            q.am = q.zf + slots;
            q.ex = q.am + slots;
            q.pair = (unsigned*)(q.ex + slots);
            q.fwd = q.pair + slots;
            q.meta = (int*)(q.fwd + slots);
            return q;
        };
        int lcap = kStretchLevels;
        if (const char* ev = std::getenv("AMH_STRETCH_LEVELS")) {      /* test switch: forces the overflow bucket */
            const int v = std::atoi(ev);
            if (v >= 2 && v <= kStretchLevels) lcap = v;
        }
        int blk = 512;                                                  /* threads of the ensemble's CTA: 512 leaves 128 registers per thread (no spills) */
        if (const char* ev = std::getenv("AMH_STRETCH_BLOCK")) {
            const int v = std::atoi(ev);
            if (v == 768 || v == 1024) blk = v;
        }
        static const char* cl_env = std::getenv("AMH_STRETCH_CLUSTER");                                     /* A/B switch: 0 / 1 */
        const bool use_cluster = T::kind != AMH_TARGET_USER && blk == 512 &&
                                 (cl_env ? std::atoi(cl_env) != 0 : (2 * nens <= r.ctx->sm_count && a.n_walkers >= 1024));
        /* forwarding slots: what is left of 200 KB of shared memory beside the per-walker flags and counters
         * (cluster variant: + one 8-byte mbarrier per slot) */
        const size_t fixed_sm = (size_t)a.n_walkers * (sizeof(unsigned) + sizeof(unsigned short) + 1) + 32;
        const size_t slot_sm = sizeof(double) * r.dim + (use_cluster ? 8 : 0);
        long long fcap = ((long long)(200 * 1024) - (long long)fixed_sm) / (long long)slot_sm;
        fcap = std::max<long long>(0, std::min<long long>(fcap, 65534));
        if (const char* ev = std::getenv("AMH_STRETCH_FWD")) fcap = std::min<long long>(fcap, std::atoll(ev));   /* test switch */
        static const bool no_ahead = std::getenv("AMH_STRETCH_NO_AHEAD") != nullptr;                              /* A/B switch */
        constexpr int PB = 512;                   /* two plan CTAs per SM: one computes while the other sits at a barrier */
        const size_t smemp = (size_t)a.n_walkers * 3 * sizeof(int);
        auto kp = stretch_plan_kernel<PB>;
        if (smemp > 40 * 1024) AMH_CUDA_TRY(cudaFuncSetAttribute(kp, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemp));
        auto enqueue_plan = [&](cudaStream_t st, int b, unsigned long long step0) -> int {
            kp<<<(unsigned)(nsteps * nens), PB, smemp, st>>>(plan_at(b), r.seeds, r.n, (int)a.n_walkers, r.dim, step0, a.a, lcap, (int)fcap);
            AMH_CUDA_TRY(cudaGetLastError());
            AMH_CUDA_TRY(cudaEventRecord(r.ev_plan[b], st));
            r.plan_step0[b] = (long long)step0;
            r.plan_nsteps[b] = nsteps;
            r.launches += 1;
            return AMH_OK;
        };
        int cur = -1;
        if (nsteps > 0) {
            for (int b = 0; b < 2; ++b)
                if (r.plan_step0[b] == (long long)a.step0 && r.plan_nsteps[b] == nsteps) cur = b;
            if (cur >= 0) {
                AMH_CUDA_TRY(cudaStreamWaitEvent(r.ctx->stream, r.ev_plan[cur], 0));      /* made ahead on the aux stream */
            } else {
                /* nothing usable was made ahead (first launch, different length, state was reset): plan now.  Buffer 0
                 * may hold a stale plan still being written by the aux stream: order behind it. */
                cur = 0;
                if (r.plan_step0[0] >= 0) AMH_CUDA_TRY(cudaStreamWaitEvent(r.ctx->stream, r.ev_plan[0], 0));
                const int rcp = enqueue_plan(r.ctx->stream, 0, a.step0);
                if (rcp) return rcp;
            }
            plan = plan_at(cur);
        } else {
            plan = plan_at(0);
        }
        const size_t smemv = (size_t)fcap * slot_sm + fixed_sm;
        if constexpr (T::kind == AMH_TARGET_USER) {
            static_assert(DMAX == 0, "RK_FLOW* name stretch_sweep_flow_kernel<0, TUser, BL>");
            int fcap_i = (int)fcap;
            void* params[] = {(void*)&a, (void*)&plan, (void*)&tp, (void*)&RA, (void*)&RB, (void*)&fcap_i};
            const int rc = rtc_launch(r, blk == 512 ? RK_FLOW512 : blk == 768 ? RK_FLOW768 : RK_FLOW1024, (unsigned)nens,
                                      (unsigned)blk, smemv, params);
            if (rc) return rc;
        } else {
            if (use_cluster) {
                auto kf2 = stretch_sweep_flow2_kernel<DMAX, T, 512>;
                if (smemv > 48 * 1024)
                    AMH_CUDA_TRY(cudaFuncSetAttribute(kf2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemv));
                kf2<<<(unsigned)(2 * nens), 512, smemv, r.ctx->stream>>>(a, plan, tp, RA, RB, (int)fcap);
            } else
#define AMH_STRETCH_LAUNCH(BL)                                                                                             \
        do {                                                                                                               \
            auto kf = stretch_sweep_flow_kernel<DMAX, T, BL>;                                                              \
            if (smemv > 48 * 1024)                                                                                         \
                AMH_CUDA_TRY(cudaFuncSetAttribute(kf, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemv));           \
            kf<<<(unsigned)nens, BL, smemv, r.ctx->stream>>>(a, plan, tp, RA, RB, (int)fcap);                              \
        } while (0)
            if (blk == 512) AMH_STRETCH_LAUNCH(512);
            else if (blk == 768) AMH_STRETCH_LAUNCH(768);
            else AMH_STRETCH_LAUNCH(1024);
#undef AMH_STRETCH_LAUNCH
            AMH_CUDA_TRY(cudaGetLastError());
        }
        r.launches += 1;
        r.pending_launches += 1;
        if (cur >= 0) {
            AMH_CUDA_TRY(cudaEventRecord(r.ev_sweep[cur], r.ctx->stream));
            if (!no_ahead) {
                /* the next launch will most likely continue with the same number of sweeps: make its plan now, on the
                 * aux stream, behind the last sweep that read the other buffer */
                const int o = cur ^ 1;
                AMH_CUDA_TRY(cudaStreamWaitEvent(r.aux_stream, r.ev_sweep[o], 0));
                const int rcp = enqueue_plan(r.aux_stream, o, a.step0 + (unsigned long long)nsteps);
                if (rcp) return rcp;
            }
        }
        return AMH_OK;
    }
    const size_t smem = (size_t)a.n_walkers * (sizeof(int) + 1) + 16;
    if (smem > 200 * 1024) return fail(AMH_ERR_UNSUPPORTED, "Ensemble on the device supports n_walkers <= 40000");
    const unsigned grid = (unsigned)(r.n / a.n_walkers);
    if constexpr (T::kind == AMH_TARGET_USER) {
        static_assert(DMAX == 0 && BLOCK == 1024, "RK_STRETCH names stretch_sweep_kernel<0, TUser, 1024>");
        void* params[] = {(void*)&a, (void*)&tp};
        const int rc = rtc_launch(r, RK_STRETCH, grid, BLOCK, smem, params);
        if (rc) return rc;
    } else {
        auto kern = stretch_sweep_kernel<DMAX, T, BLOCK>;
        if (smem > 48 * 1024) AMH_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<grid, BLOCK, smem, r.ctx->stream>>>(a, tp);
        AMH_CUDA_TRY(cudaGetLastError());
    }
    if (nsteps & 1) {          /* the current state now lives in the other buffer */
        std::swap(r.X, r.X2);
        std::swap(r.lp, r.lp2);
    }
    r.launches += 1;
    r.pending_launches += 1;
    return AMH_OK;
}

#ifndef AMH_STRETCH_EXTRA_TU
/* Sweeps per launch when the caller does not say: 16, and for the K2R shapes as many of 64 / 32 as keep the two plan
 * buffers under 2 GiB -- a launch ends with the deepest dependency chain of its slowest ensemble while the other SMs
 * idle, so longer launches pay: config 3 at 8 / 16 / 32 / 64 sweeps per launch 1.05 / 1.13 / 1.15 / 1.18e10 moves/s
 * (profiles/r2_c3_shape_sweep.txt); 64 sweeps of 64 x 4 096 walkers are 2 x 0.8 GB of plan entries. */
int stretch_default_steps_per_launch(const amh_run& r) {
    const long long nw = r.sampler->d.n_walkers;
    if (nw < 1 || r.n % nw) return 16;
    const long long nens = r.n / nw;
    if (!stretch_res_shape(nw, nens, r.ctx->sm_count) || std::getenv("AMH_STRETCH_RES")) return 16;
    const double per_sweep = (double)nens * 2.0 * ((double)res_nwp(nw) * 4 * sizeof(double) + kResMeta * sizeof(int));
    for (int spl : {64, 32})
        if (2.0 * spl * per_sweep <= 2147483648.0) return spl;
    return 16;
}

template <class T>
int launch_stretch_dim(amh_run& r, int nsteps, const SaveArgs& sv) {
    switch (r.dim) {          /* exact-dimension instantiations; more of them in amh_launch_stretch_dims.cu; everything else is generic */
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
    {   /* the second translation unit's exact-dimension instantiations (compiled in parallel with this one) */
        bool taken = false;
        const int rc = launch_stretch_more_dims(r, nsteps, sv, taken);
        if (taken) return rc;
    }
    switch (r.target->kind) {
    case AMH_TARGET_MVNORMAL: return launch_stretch_dim<TMvNormal>(r, nsteps, sv);
    case AMH_TARGET_GAUSS_PREC: return launch_stretch_dim<TGaussPrec>(r, nsteps, sv);
    case AMH_TARGET_ROSENBROCK: return launch_stretch_dim<TRosenbrock>(r, nsteps, sv);
    case AMH_TARGET_IID_NORMAL: return launch_stretch_t<2, TIidNormal>(r, nsteps, sv);
    case AMH_TARGET_NIG_TOY:
    case AMH_TARGET_NIG_TOY_LOG: return launch_stretch_t<2, TNig>(r, nsteps, sv);
    case AMH_TARGET_LOGISTIC: return launch_stretch_t<0, TLogistic>(r, nsteps, sv);
    case AMH_TARGET_USER: return launch_stretch_t<0, TUser>(r, nsteps, sv);
    }
    return fail(AMH_ERR_INVALID, "unknown target kind");
}
#endif  /* AMH_STRETCH_EXTRA_TU */

}  // namespace amhh
