/* amh_launch_stretch_res.cuh -- K2R: the exact sequential stretch sweep (emcee.jl:39-102) with the whole ensemble
 * RESIDENT in the shared memory of a 2-CTA cluster for all sweeps of a launch.
 *
 * Included by amh_launch_stretch.cu (after the @rtc region: the run-time-compiled user targets keep K2F).
 *
 * Why: K2F keeps the walkers in L2.  Its profile (profiles/r2_k2f_mbar_ncu_summary.txt) is a chain of exposed round
 * trips per 32-move chunk -- own record from L2 (~740 cycles), partner record, record store, remote counters -- with
 * only 16 warps per SM to hide them: ~9 000 cycles per chunk, 4 chunks per warp and sweep.  Here nothing of a sweep
 * leaves the two SMs except the (prefetched) plan.
 *
 * Layout.  CTA r of the cluster OWNS the walkers of parity r (i & 1 == r; parity, not halves: low-index walkers are
 * mostly level 0, so halves would put all the deep levels on one SM): their records [x_0..x_{d-1}, lp] live in its
 * shared memory, are updated IN PLACE and are moved only by its threads.  4 096 walkers x 11 doubles = 352 KB do not
 * fit twice, so there is no old/new double buffer; the sequential semantics (emcee.jl:53: walker i reads the NEW
 * value of partner idx < i and the OLD value of idx > i) is kept by ordering instead:
 *
 *   level 0  = the walkers with idx > i: the ONLY readers of old values.  They run first, bulk-synchronously, in
 *              WINDOWS of ascending walker index (at most BLOCK walkers per CTA and window): read own + partner,
 *              barrier.cluster.arrive, compute, barrier.cluster.wait, write in place.  A window reads only walkers of
 *              the same or a later window (idx > i) and all its reads are complete before any of its writes, so no old
 *              value is overwritten before its last reader has seen it; the barrier latency hides behind the compute.
 *   level>=1 = dataflow, as in K2F: 32-slot chunks in level order per CTA, a lane waits for its partner's new value.
 *              Their partners (idx < i) were all written in this sweep; their own in-place write can hurt nobody
 *              (old-value readers are level 0 and done).  Hand-off of a new value:
 *                same CTA:   record in place, __threadfence_block, version flag (shared memory, CTA scope);
 *                other CTA:  the mover PUSHES the value into a forwarding slot of the reader's CTA with st.async, which
 *                            completes an mbarrier there (no cluster-scope fence: those are MEMBAR.ALL.GPU + an L1
 *                            invalidation); the reader waits on a barrier of its own CTA and reads its own shared memory.
 *                            The plan knows which walkers have a reader in the other CTA (a few hundred per sweep);
 *              level-0 partners need no hand-off at all: the barrier that closes level 0 orders them (the other CTA's
 *              are read through distributed shared memory).
 *   overflow = levels >= lcap - 1 (never reached in practice, forced by tests): one thread, ascending walker order.
 */
#pragma once

namespace amhh {

constexpr int kResMeta = 64;          /* ints of meta data per (sweep, ensemble, cta) */
constexpr int kResMaxWin = 48;        /* level-0 windows per sweep: meta[8 .. 8 + kResMaxWin] = window starts */
constexpr unsigned kResNoSlot = 0xfffu;

/* info word of a plan entry:  bits 0-11 forwarding slot of the PARTNER's value in this CTA (reader kind 2),
 * 12-23 forwarding slot (in the other CTA) this walker's new value is pushed to, or kResNoSlot,
 * 24-25 reader kind: 0 = partner value needs no wait (old value, or a level-0 partner), 1 = same-CTA version flag,
 *                   2 = pushed into this CTA (slot in bits 0-11), 3 = other CTA, no slot left: remote flag + remote read,
 * 26 = a later walker of this CTA reads the new value in this sweep (raise the version flag),
 * 27 = a later walker of the OTHER CTA reads it and there was no slot left (cluster-scope release before the flag) */
struct StretchPlanR {
    double* ent;               /* [nsteps][n_ensembles][2][nwp][4] one 32-byte entry per slot: bits(self | partner << 16 | info << 32)
                                  (self | partner = the sentinel in padding slots), z, (d-1) log z, the exponential */
    int* meta;                 /* [nsteps][n_ensembles][2][kResMeta]: 0 level-0 walkers, 1 windows, 2/3 dataflow slots [lo, hi),
                                  4/5 overflow slots [lo, hi), 6 forwarding slots of this CTA in use, 8.. window starts */
    long long nwp;             /* slots per (sweep, ensemble, cta) */
};

struct ResEntry { unsigned pair, info; double z, am, ex; };
__device__ __forceinline__ ResEntry res_entry_load(const double* __restrict__ ent, size_t slot) {
    double e0, e1, e2, e3;
    ld256(ent + slot * 4, e0, e1, e2, e3);
    const unsigned long long b = (unsigned long long)__double_as_longlong(e0);
    ResEntry r;
    r.pair = (unsigned)b; r.info = (unsigned)(b >> 32); r.z = e1; r.am = e2; r.ex = e3;
    return r;
}

/* the shapes K2R is chosen for: few, large ensembles (two SMs each) */
inline bool stretch_res_shape(long long nw, long long nens, int sm_count) { return 2 * nens <= sm_count && nw >= 1024 && nw <= 16384; }

__host__ __device__ inline long long res_nwp(long long nw) { return (((nw + 1) / 2 + 31) & ~31ll) + 32 * (kStretchLevels + 1); }

template <int BLOCK>
__global__ void __launch_bounds__(BLOCK)
stretch_plan_res_kernel(StretchPlanR o, const unsigned long long* __restrict__ seeds, long long n, int nw, int d,
                        unsigned long long step0, double aa, int lcap, int fcap, int wincap) {
    extern __shared__ __align__(16) unsigned char smem_pr[];
    double* ubuf = reinterpret_cast<double*>(smem_pr);                    /* [nw] the stretch uniform */
    int* partner = reinterpret_cast<int*>(ubuf + nw);                     /* [nw] */
    int* slotof = partner + nw;                                           /* [nw] slot in the owner's list */
    unsigned* r01 = reinterpret_cast<unsigned*>(slotof + nw);             /* [nw + 1] level-0 walkers of parity 0 | 1 << 16 below index i */
    unsigned short* push = reinterpret_cast<unsigned short*>(r01 + nw + 1);   /* [nw] 0/1 flag, later the slot */
    unsigned char* lvl = reinterpret_cast<unsigned char*>(push + nw);     /* [nw] */
    unsigned char* wl = lvl + nw;                                         /* [nw] same-CTA reader */
    __shared__ int hist[2][kStretchLevels];
    __shared__ int start[2][kStretchLevels + 1];
    __shared__ unsigned long long wsum[BLOCK / 32];
    __shared__ int s_nwin, s_win[kResMaxWin + 1];
    const int tid = threadIdx.x;
    const long long nens = n / nw;
    const int s = (int)(blockIdx.x / nens);
    const long long en = blockIdx.x % nens;
    const unsigned long long k = step0 + (unsigned long long)s + 1ull;
    const unsigned long long seed = seeds[en];
    if (tid < 2 * kStretchLevels) hist[tid / kStretchLevels][tid % kStretchLevels] = 0;
    for (int i = tid; i < nw; i += BLOCK) {
        const unsigned long long blk = (k * (unsigned long long)nw + (unsigned long long)i) * 2ull;
        const amh::Block b0 = amh::stream_block(seed, blk, 0u);
        /* idx = mod1(i + rand(1:(n-1)), n)  (emcee.jl:52) */
        const long long rr = (long long)amh::bounded(b0.v[0], b0.v[1], (unsigned long long)(nw - 1));
        int pj = i + (int)rr + 1;
        if (pj >= nw) pj -= nw;
        partner[i] = pj;
        ubuf[i] = amh::u01(b0.v[2], b0.v[3]);
        push[i] = 0;
        wl[i] = 0;
    }
    __syncthreads();
    for (int i = tid; i < nw; i += BLOCK) {
        int lv = 0, cur = i, j = partner[i];
        while (j < cur && lv < lcap - 1) { ++lv; cur = j; j = partner[cur]; }
        lvl[i] = (unsigned char)lv;
    }
    __syncthreads();
    /* who has to hand its new value to whom: only partners of level >= 1 (level 0 is closed by a barrier) and only
     * readers outside the overflow bucket (it runs after everything else) */
    for (int i = tid; i < nw; i += BLOCK) {
        const int j = partner[i];
        if (j < i && lvl[j] >= 1 && lvl[i] < lcap - 1) {
            if (((i ^ j) & 1) == 0) wl[j] = 1; else push[j] = 1;          /* benign races: everybody writes 1 */
        }
    }
    __syncthreads();
    /* one block-wide exclusive scan of four counters (16 bits each): level-0 walkers of parity 0 / 1, pushes into CTA 0 / 1.
     * Thread t owns the contiguous walkers [t * per, (t + 1) * per): ranks ascend with the walker index. */
    {
        const int per = (nw + BLOCK - 1) / BLOCK;
        const int lo = min(nw, tid * per), hi = min(nw, lo + per);
        unsigned long long cnt = 0;
        for (int i = lo; i < hi; ++i) {
            if (lvl[i] == 0) cnt += 1ull << (16 * (i & 1));
            if (push[i]) cnt += 1ull << (32 + 16 * ((i & 1) ^ 1));         /* destination = the other CTA */
        }
        unsigned long long inc = cnt;
        for (int o2 = 1; o2 < 32; o2 <<= 1) {
            const unsigned long long v = __shfl_up_sync(0xffffffffu, inc, o2);
            if ((tid & 31) >= o2) inc += v;
        }
        if ((tid & 31) == 31) wsum[tid >> 5] = inc;
        __syncthreads();
        if (tid < 32) {
            unsigned long long v = (tid < BLOCK / 32) ? wsum[tid] : 0ull;
            for (int o2 = 1; o2 < 32; o2 <<= 1) {
                const unsigned long long u = __shfl_up_sync(0xffffffffu, v, o2);
                if (tid >= o2) v += u;
            }
            if (tid < BLOCK / 32) wsum[tid] = v;                           /* inclusive over warps */
        }
        __syncthreads();
        unsigned long long run = inc - cnt + ((tid >> 5) ? wsum[(tid >> 5) - 1] : 0ull);
        for (int i = lo; i < hi; ++i) {
            const int p = i & 1;
            r01[i] = (unsigned)(run & 0xffffffffull);
            if (lvl[i] == 0) {
                slotof[i] = (int)((run >> (16 * p)) & 0xffffull);          /* level 0: ascending walker order */
                run += 1ull << (16 * p);
            }
            if (push[i]) {
                const int q = (int)((run >> (32 + 16 * (p ^ 1))) & 0xffffull);
                run += 1ull << (32 + 16 * (p ^ 1));
                push[i] = (unsigned short)(q < fcap ? q : 0xfffe);         /* 0xfffe: no slot left */
            } else {
                push[i] = 0xffff;
            }
        }
        if (tid == BLOCK - 1) r01[nw] = (unsigned)(run & 0xffffffffull);
    }
    /* levels >= 1: rank within (cta, level) in any order */
    for (int i = tid; i < nw; i += BLOCK)
        if (lvl[i] >= 1) slotof[i] = atomicAdd(&hist[i & 1][lvl[i]], 1);
    __syncthreads();
    const unsigned tot = (unsigned)(wsum[BLOCK / 32 - 1] & 0xffffffffull);
    const unsigned long long totp = wsum[BLOCK / 32 - 1] >> 32;
    if (tid < 2) {
        const int p = tid;
        const int n0 = (int)((tot >> (16 * p)) & 0xffffu);
        int acc = (n0 + 31) & ~31;
        start[p][0] = 0;
        for (int l = 1; l < kStretchLevels; ++l) { start[p][l] = acc; acc += (hist[p][l] + 31) & ~31; }
        start[p][kStretchLevels] = acc;
    }
    if (tid == 32) {
        /* level-0 windows: consecutive ranges of the walker index with at most `wincap` level-0 walkers of either parity */
        int nwin = 0, b = 0;
        s_win[0] = 0;
        while (b < nw && nwin < kResMaxWin) {
            const unsigned rb = r01[b];
            int lo = b + 1, hi = nw;                                       /* largest e with both counts in [b, e) <= wincap */
            while (lo < hi) {
                const int mid = (lo + hi + 1) >> 1;
                const unsigned rm = r01[mid];
                const bool ok = (int)((rm & 0xffffu) - (rb & 0xffffu)) <= wincap && (int)((rm >> 16) - (rb >> 16)) <= wincap;
                if (ok) lo = mid; else hi = mid - 1;
            }
            b = lo;
            s_win[++nwin] = b;
        }
        if (b < nw) { s_win[nwin] = nw; nwin = -1; }                       /* more windows than the meta block holds: flagged */
        s_nwin = nwin;
    }
    __syncthreads();
    const size_t blk_se = (size_t)s * nens + en;
    if (tid < 2) {
        const int p = tid;
        int* m = o.meta + (blk_se * 2 + p) * kResMeta;
        const int n0 = (int)((tot >> (16 * p)) & 0xffffu);
        m[0] = n0;
        m[1] = s_nwin;
        m[2] = start[p][1];
        m[3] = start[p][lcap - 1];
        m[4] = start[p][lcap - 1];
        m[5] = start[p][lcap - 1] + hist[p][lcap - 1];
        m[6] = min((int)((totp >> (16 * p)) & 0xffffull), fcap);
        m[7] = 0;
        const int nwin = s_nwin < 0 ? 0 : s_nwin;
        for (int w = 0; w <= nwin; ++w) m[8 + w] = (int)((r01[s_win[w]] >> (16 * p)) & 0xffffu);
    }
    /* sentinels in the padding of every level of both lists */
    for (int pl = tid >> 5; pl < 2 * kStretchLevels; pl += BLOCK >> 5) {
        const int p = pl / kStretchLevels, l = pl % kStretchLevels;
        const int cntl = l == 0 ? (int)((tot >> (16 * p)) & 0xffffu) : hist[p][l];
        const int q = start[p][l] + cntl + (tid & 31);
        if (q < start[p][l + 1])
            st256(o.ent + ((blk_se * 2 + p) * (size_t)o.nwp + q) * 4, __longlong_as_double((long long)kStretchSentinel), 0.0, 0.0, 0.0);
    }
    /* z = t^2 / a for every move: one reciprocal refinement of `a` per thread + Markstein's correction (amh_fastmath.cuh;
     * t^2 is in [1, a^2], so only `a` needs the range check) */
    const bool zfast = fast_div_ok(1.0, aa) && fast_div_ok(aa * aa, aa);
    const double raa = zfast ? rcp_refined(aa) : 0.0;
    for (int i = tid; i < nw; i += BLOCK) {
        const unsigned long long blk = (k * (unsigned long long)nw + (unsigned long long)i) * 2ull;
        const amh::Block b1 = amh::stream_block(seed, blk + 1ull, 0u);
        const int p = i & 1, lv = lvl[i];
        const size_t slot = (blk_se * 2 + p) * (size_t)o.nwp + start[p][lv] + slotof[i];
        const double tt = (aa - 1.0) * ubuf[i] + 1.0;
        const double t2 = tt * tt;
        const double z = zfast ? div_with_rcp(t2, aa, raa) : t2 / aa;   /* the correctly rounded quotient either way */
        const int pj = partner[i];
        unsigned rk = 0u, pslot = 0u;
        if (pj < i && lv < lcap - 1 && lvl[pj] >= 1) {
            if (((i ^ pj) & 1) == 0) rk = 1u;
            else if (push[pj] < 0xfffe) { rk = 2u; pslot = push[pj]; }
            else rk = 3u;
        }
        const unsigned sslot = (lv >= 1 && push[i] < 0xfffe) ? push[i] : kResNoSlot;
        const unsigned fl = (lv >= 1 && wl[i] ? 1u << 26 : 0u) | (lv >= 1 && push[i] == 0xfffe ? 1u << 27 : 0u);
        const unsigned long long w0 = (unsigned long long)((unsigned)i | ((unsigned)pj << 16)) |
                                      ((unsigned long long)(pslot | (sslot << 12) | (rk << 24) | fl) << 32);
        st256(o.ent + slot * 4, __longlong_as_double((long long)w0), z, (double)(d - 1) * amh::log_(z), amh::exponential(b1.v[0], b1.v[1]));
    }
}

struct ResSmem {
    double* rec;                          /* [nwl][d + 1] records of the walkers this CTA owns (local index i >> 1) */
    double* fwd;                          /* [fcap][d] values pushed in by the other CTA */
    unsigned bar;                         /* shared::cta address of the slot barriers [fcap] */
    unsigned bar_r, fwd_r;                /* the other CTA's, shared::cluster addresses */
    unsigned rec_r;                       /* the other CTA's records, shared::cluster address */
    unsigned ver_r;                       /* the other CTA's version flags */
    double* G;                            /* this ensemble's records in global memory (L2): the copy the OTHER CTA reads whenever
                                             a cluster barrier orders the access (old values, level-0 partners): random 8-byte
                                             reads through distributed shared memory are bound by the request rate of the
                                             SM-to-SM network (~3 cycles each, profiles/r2_k2r_first_ncu_summary.txt) */
    volatile unsigned short* ver;         /* [nwl] last sweep (1-based) the walker's new value was published */
    unsigned* nacc;                       /* [nwl] accepted moves of this launch | last accept flag << 31 */
};

__device__ __forceinline__ double ld_cluster_f64(unsigned addr) {
    double v;
    asm volatile("ld.shared::cluster.f64 %0, [%1];" : "=d"(v) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ unsigned short ld_acquire_cluster_remote_u16(unsigned addr) {
    unsigned short v;
    asm volatile("ld.acquire.cluster.shared::cluster.u16 %0, [%1];" : "=h"(v) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }

__host__ __device__ constexpr int res_gs(int dmax) { return (dmax + 1 + 3) & ~3; }      /* doubles per record of the global mirror (32-byte multiples) */

/* partner coordinates into o[]: own CTA's record from shared memory; the other CTA's from the global mirror (`dsmem`:
 * through distributed shared memory instead -- the rare paths that no barrier orders) */
template <int DMAX>
__device__ __forceinline__ void res_load_partner(double (&o)[DMAX], const ResSmem& sm, unsigned rank, int idx, bool dsmem = false) {
    if (((unsigned)idx & 1u) == rank) {
        const double* p = sm.rec + (size_t)(idx >> 1) * (DMAX + 1);
#pragma unroll
        for (int j = 0; j < DMAX; ++j) o[j] = p[j];
    } else if (dsmem) {
        const unsigned a = sm.rec_r + (unsigned)(idx >> 1) * (unsigned)((DMAX + 1) * 8);
#pragma unroll
        for (int j = 0; j < DMAX; ++j) o[j] = ld_cluster_f64(a + j * 8u);
    } else {
        constexpr int GS = res_gs(DMAX);
        const double* g = sm.G + (size_t)idx * GS;
        double t[GS];
#pragma unroll
        for (int j = 0; j < ((DMAX + 3) & ~3); j += 4) ld256(g + j, t[j], t[j + 1], t[j + 2], t[j + 3]);
#pragma unroll
        for (int j = 0; j < DMAX; ++j) o[j] = t[j];
    }
}

/* an accepted move: the record in place ... */
template <int DMAX>
__device__ __forceinline__ void res_store_record(const ResSmem& sm, int i, const double (&w)[DMAX + 1]) {
    double* rec = sm.rec + (size_t)(i >> 1) * (DMAX + 1);
#pragma unroll
    for (int j = 0; j < DMAX + 1; ++j) rec[j] = w[j];
}
/* ... and its copy in the global mirror (after the hand-off: nobody reads it before the next cluster barrier) */
template <int DMAX>
__device__ __forceinline__ void res_store_mirror(const ResSmem& sm, int i, const double (&w)[DMAX + 1]) {
    constexpr int GS = res_gs(DMAX);
    double t[GS];
#pragma unroll
    for (int j = 0; j < GS; ++j) t[j] = j <= DMAX ? w[j < DMAX + 1 ? j : DMAX] : 0.0;
    double* g = sm.G + (size_t)i * GS;
#pragma unroll
    for (int j = 0; j < GS; j += 4) st256(g + j, t[j], t[j + 1], t[j + 2], t[j + 3]);
}
__device__ __forceinline__ void fence_cta() { asm volatile("fence.acq_rel.cta;" ::: "memory"); }

/* the move itself (emcee.jl:70-102): w = own record (in: current, out: new), o = partner; returns the accept flag */
template <int DMAX, class T>
__device__ __forceinline__ bool res_move(const typename T::template Params<DMAX>& tp, double (&w)[DMAX + 1], const double (&o)[DMAX],
                                         double z, double am, double ex) {
    double y[DMAX];
#pragma unroll
    for (int j = 0; j < DMAX; ++j) y[j] = o[j] + z * (w[j] - o[j]);
    const double lpw = w[DMAX];
    const double lpy = T::template logp<DMAX>(y, DMAX, tp);
    const double alpha = (am + lpy) - lpw;
    const bool acc = (-ex <= alpha);                         /* emcee.jl:93 (non-strict) */
#pragma unroll
    for (int j = 0; j < DMAX; ++j) w[j] = acc ? y[j] : w[j];
    w[DMAX] = acc ? lpy : lpw;
    return acc;
}

template <int DMAX, class T, int BLOCK>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(BLOCK, 1)
stretch_sweep_res_kernel(const __grid_constant__ StretchArgs a, const __grid_constant__ StretchPlanR plan,
                         const __grid_constant__ typename T::template Params<DMAX> tp, double* Gall, int fcap) {
    static_assert(DMAX > 0, "K2R has exact-dimension instantiations only");
    constexpr int RS = DMAX + 1;
    constexpr int GS = res_gs(DMAX);
    constexpr int NWARP = BLOCK / 32;
    extern __shared__ __align__(16) double smem_rs[];
    const int nw = (int)a.n_walkers;
    const int nwl = (nw + 1) >> 1;
    const unsigned rank = cluster_ctarank();
    const int nown = (nw + 1 - (int)rank) >> 1;                /* walkers of parity `rank` */
    ResSmem sm;
    sm.rec = smem_rs;
    sm.fwd = sm.rec + (((size_t)nwl * RS + 1) & ~(size_t)1);          /* 16-byte aligned: st.async.v2 */
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(sm.fwd + (size_t)fcap * DMAX);
    sm.nacc = reinterpret_cast<unsigned*>(bars + fcap);
    unsigned short* ver_nv = reinterpret_cast<unsigned short*>(sm.nacc + nwl);
    sm.ver = ver_nv;
    sm.bar = smem_u32(bars);
    sm.bar_r = map_cta_u32(sm.bar, rank ^ 1u);
    sm.fwd_r = map_cta_u32(smem_u32(sm.fwd), rank ^ 1u);
    sm.rec_r = map_cta_u32(smem_u32(sm.rec), rank ^ 1u);
    sm.ver_r = map_cta_u32(smem_u32(ver_nv), rank ^ 1u);
    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const long long en = blockIdx.x >> 1;
    const long long nens = a.st.n / nw;
    const long long base = en * nw;
    const size_t pl_stride = (size_t)plan.nwp;
    sm.G = Gall + (size_t)base * GS;
    /* prologue: [dim][chain] state -> records of the own walkers */
    int fmax = 0;
    for (int s = 0; s < a.nsteps; ++s) fmax = max(fmax, plan.meta[(((size_t)s * nens + en) * 2 + rank) * kResMeta + 6]);
    for (int q = tid; q < fmax; q += BLOCK) mbar_init(sm.bar + q * 8u, 1u);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    for (int li = tid; li < nown; li += BLOCK) {
        const long long ch = base + 2 * li + (int)rank;
        double* rec = sm.rec + (size_t)li * RS;
#pragma unroll
        for (int j = 0; j < DMAX; ++j) rec[j] = a.st.X[(long long)j * a.st.pitch + ch];
        rec[DMAX] = a.st.lp[ch];
        ver_nv[li] = 0;
        sm.nacc[li] = (unsigned)a.st.acc[ch] << 31;
        double w0[RS];
#pragma unroll
        for (int j = 0; j < RS; ++j) w0[j] = rec[j];
        res_store_mirror<DMAX>(sm, 2 * li + (int)rank, w0);
    }
    /* what a sweep needs before it can start, fetched one sweep ahead (three dependent L2 round trips otherwise) */
    struct SweepHead {
        int nwin, dlo, dhi, olo, ohi, nfwd, olo2, ohi2, w0, w1, w2;
        ResEntry e_w, e_w2;                                     /* this thread's entries of the first two windows */
    };
    auto no_entry = [] { ResEntry x; x.pair = kStretchSentinel; x.info = 0u; x.z = x.am = x.ex = 0.0; return x; };
    auto load_head = [&](int s) {
        SweepHead h;
        const size_t pse = ((size_t)s * nens + en) * 2 + rank;
        const int* __restrict__ m = plan.meta + pse * kResMeta;
        const int* __restrict__ m2 = plan.meta + (pse ^ 1) * kResMeta;
        h.nwin = m[1]; h.dlo = m[2]; h.dhi = m[3]; h.olo = m[4]; h.ohi = m[5]; h.nfwd = m[6];
        h.olo2 = m2[4]; h.ohi2 = m2[5];
        h.w0 = m[8]; h.w1 = m[9]; h.w2 = m[10];
        const double* __restrict__ ent = plan.ent + pse * pl_stride * 4;
        h.e_w = h.e_w2 = no_entry();
        if (h.nwin > 0 && h.w0 + tid < h.w1) h.e_w = res_entry_load(ent, (size_t)(h.w0 + tid));
        if (h.nwin > 1 && h.w1 + tid < h.w2) h.e_w2 = res_entry_load(ent, (size_t)(h.w1 + tid));
        return h;
    };
    SweepHead hd;
    if (a.nsteps > 0) hd = load_head(0);
    cluster_arrive(); cluster_wait();

    for (int s = 0; s < a.nsteps; ++s) {
        const size_t pse = ((size_t)s * nens + en) * 2 + rank;
        const int* __restrict__ meta = plan.meta + pse * kResMeta;
        const double* __restrict__ ent = plan.ent + pse * pl_stride * 4;
        const unsigned short want = (unsigned short)(s + 1);
        const unsigned par = (unsigned)s & 1u;
        const int nwin = hd.nwin, dhi = hd.dhi, olo = hd.olo, ohi = hd.ohi, nfwd = hd.nfwd, olo2 = hd.olo2, ohi2 = hd.ohi2;
        if (nwin < 0) __trap();                                  /* more level-0 windows than the plan can describe: the host sizes them so that this cannot happen */
        int q = hd.dlo + warp * 32 + lane;
        ResEntry e = no_entry(), e2 = no_entry();
        /* ---- level 0: windows of ascending walker index; reads | barrier | writes.  Everything that comes from L2 is
         * fetched ahead: plan entries two windows ahead, the other CTA's partners (global mirror) one window ahead (a
         * window only reads walkers that no earlier window writes), and the mirror copy of a window's accepted moves is
         * stored after the NEXT window's arrive, so that no arrive waits for a store acknowledgement from L2 ---- */
        {
            ResEntry ew = hd.e_w, ew2 = hd.e_w2;
            int w2 = hd.w2;                                     /* end of window w + 1 */
            double og[DMAX];                                    /* partner of the current window's move when it lives in the other CTA */
            auto fetch_remote = [&](const ResEntry& x) {
                if (x.pair != kStretchSentinel && (((x.pair >> 16) & 1u) != rank)) {
                    const double* g = sm.G + (size_t)(x.pair >> 16) * GS;
                    double t[GS];
#pragma unroll
                    for (int j = 0; j < ((DMAX + 3) & ~3); j += 4) ld256(g + j, t[j], t[j + 1], t[j + 2], t[j + 3]);
#pragma unroll
                    for (int j = 0; j < DMAX; ++j) og[j] = t[j];
                }
            };
            fetch_remote(ew);
            for (int w = 0; w < nwin; ++w) {
                const ResEntry ec = ew;
                const bool have = ec.pair != kStretchSentinel;
                const int i0 = (int)(ec.pair & 0xffffu), idx0 = (int)(ec.pair >> 16), li = i0 >> 1;
                double wr[RS], o[DMAX];
                if (have) {
                    const double* rec = sm.rec + (size_t)li * RS;
#pragma unroll
                    for (int j = 0; j < RS; ++j) wr[j] = rec[j];
                    if (((unsigned)idx0 & 1u) == rank) {                            /* the OLD value (idx > i) */
                        const double* pr = sm.rec + (size_t)(idx0 >> 1) * RS;
#pragma unroll
                        for (int j = 0; j < DMAX; ++j) o[j] = pr[j];
                    } else {
#pragma unroll
                        for (int j = 0; j < DMAX; ++j) o[j] = og[j];
                    }
                }
                cluster_arrive();                                                   /* my reads of this window are done */
                /* next window: its entry was loaded a window ago, its remote partner starts its trip now */
                ew = ew2;
                fetch_remote(ew);
                ew2 = no_entry();
                const int w3 = (w + 2 < nwin) ? meta[11 + w] : w2;
                if (w + 2 < nwin && w2 + tid < w3) ew2 = res_entry_load(ent, (size_t)(w2 + tid));
                w2 = w3;
                bool acc = false;
                if (have) acc = res_move<DMAX, T>(tp, wr, o, ec.z, ec.am, ec.ex);
                cluster_wait();                                                     /* everybody's are */
                if (have) {
                    if (acc) { res_store_record<DMAX>(sm, i0, wr); res_store_mirror<DMAX>(sm, i0, wr); }
                    const unsigned c = sm.nacc[li] & 0x7fffffffu;
                    sm.nacc[li] = acc ? ((c + 1u) | 0x80000000u) : c;
                }
            }
            cluster_arrive();
            /* the first two dataflow chunks' entries of this lane while the barrier completes */
            if (q < dhi) e = res_entry_load(ent, (size_t)q);
            if (q + NWARP * 32 < dhi) e2 = res_entry_load(ent, (size_t)(q + NWARP * 32));
            cluster_wait();                                                         /* level 0 is written, cluster wide */
        }
        /* ---- levels >= 1: dataflow ---- */
#pragma unroll 1
        for (; q - lane < dhi; q += NWARP * 32) {
            const ResEntry ec = e;
            e = e2;                                                              /* loaded an iteration ago */
            const int qn = q + 2 * NWARP * 32;
            e2.pair = kStretchSentinel;
            if (qn < dhi) e2 = res_entry_load(ent, (size_t)qn);
            if (ec.pair != kStretchSentinel) {
                const int i = (int)(ec.pair & 0xffffu), idx = (int)(ec.pair >> 16);
                const int li = i >> 1;
                const unsigned inf_c = ec.info;
                const unsigned rk = (inf_c >> 24) & 3u;
                double wr[RS], o[DMAX];
                double* rec = sm.rec + (size_t)li * RS;
#pragma unroll
                for (int j = 0; j < RS; ++j) wr[j] = rec[j];
                if (rk == 2u) {                                                     /* pushed into this CTA */
                    const unsigned ps = inf_c & 0xfffu;
                    mbar_wait(sm.bar + ps * 8u, par);
                    const double* pf = sm.fwd + (size_t)ps * DMAX;
#pragma unroll
                    for (int j = 0; j < DMAX; ++j) o[j] = pf[j];
                } else {
                    if (rk == 1u) {                                                 /* same CTA: version flag */
                        while (sm.ver[idx >> 1] != want) { }
                        fence_cta();
                    } else if (rk == 3u) {                                          /* other CTA, no slot: its flag, remotely */
                        while (ld_acquire_cluster_remote_u16(sm.ver_r + (unsigned)(idx >> 1) * 2u) != want) { }
                    }
                    res_load_partner<DMAX>(o, sm, rank, idx, rk == 3u);
                }
                const bool acc = res_move<DMAX, T>(tp, wr, o, ec.z, ec.am, ec.ex);
                const unsigned ss = (inf_c >> 12) & 0xfffu;
                if (ss != kResNoSlot) {                                             /* a reader in the other CTA */
                    const unsigned rb = sm.bar_r + ss * 8u, rf = sm.fwd_r + ss * (unsigned)(DMAX * 8);
                    mbar_arrive_expect_tx_peer(rb, (unsigned)(DMAX * 8));
                    if constexpr (DMAX % 2 == 0) {
#pragma unroll
                        for (int j = 0; j < DMAX; j += 2) st_async_peer2(rf + j * 8u, wr[j], wr[j + 1], rb);
                    } else {
#pragma unroll
                        for (int j = 0; j < DMAX; ++j) st_async_peer(rf + j * 8u, wr[j], rb);
                    }
                }
                if (acc) res_store_record<DMAX>(sm, i, wr);
                if (inf_c & (3u << 26)) {
                    if (inf_c & (1u << 27)) fence_cluster(); else fence_cta();
                    sm.ver[li] = want;
                }
                if (acc) res_store_mirror<DMAX>(sm, i, wr);
                const unsigned c = sm.nacc[li] & 0x7fffffffu;
                sm.nacc[li] = acc ? ((c + 1u) | 0x80000000u) : c;
            }
        }
        /* the next sweep's head while this one drains */
        if (s + 1 < a.nsteps) hd = load_head(s + 1);
        /* every slot barrier of this CTA completes exactly one phase per sweep: used slots by their mover in the other
         * CTA (wait: its async stores may still be in flight), unused ones here */
        for (int qs = tid; qs < fmax; qs += BLOCK) {
            if (qs >= nfwd) mbar_arrive(sm.bar + qs * 8u);       /* unused in this sweep: completes the phase at once */
            mbar_wait(sm.bar + qs * 8u, par);
        }
        cluster_arrive(); cluster_wait();
        /* overflow bucket: one thread, increasing walker order over both CTAs' lists = the reference's own loop */
        if (ohi > olo || ohi2 > olo2) {
            if (tid == 0 && rank == 0u) {
                const double* __restrict__ ent2 = plan.ent + (pse ^ 1) * pl_stride * 4;
                int last = -1;
                for (int c = 0; c < (ohi - olo) + (ohi2 - olo2); ++c) {
                    int best = 0x7fffffff;
                    ResEntry be;
                    for (int qq = olo; qq < ohi; ++qq) {
                        const ResEntry x = res_entry_load(ent, (size_t)qq);
                        const int self = (int)(x.pair & 0xffffu);
                        if (self > last && self < best) { best = self; be = x; }
                    }
                    for (int qq = olo2; qq < ohi2; ++qq) {
                        const ResEntry x = res_entry_load(ent2, (size_t)qq);
                        const int self = (int)(x.pair & 0xffffu);
                        if (self > last && self < best) { best = self; be = x; }
                    }
                    const int i = best, idx = (int)(be.pair >> 16);
                    double* recg = map_cta(sm.rec, (unsigned)(i & 1)) + (size_t)(i >> 1) * RS;      /* generic, either CTA */
                    double wr[RS], o[DMAX];
#pragma unroll
                    for (int j = 0; j < RS; ++j) wr[j] = recg[j];
                    res_load_partner<DMAX>(o, sm, 0u, idx, true);
                    const bool acc = res_move<DMAX, T>(tp, wr, o, be.z, be.am, be.ex);
                    if (acc) {
#pragma unroll
                        for (int j = 0; j < RS; ++j) recg[j] = wr[j];
                        res_store_mirror<DMAX>(sm, i, wr);
                    }
                    unsigned* na = map_cta(sm.nacc, (unsigned)(i & 1)) + (i >> 1);
                    const unsigned cc = *na & 0x7fffffffu;
                    *na = acc ? ((cc + 1u) | 0x80000000u) : cc;
                    last = best;
                }
            }
            cluster_arrive(); cluster_wait();
        }
    }
    /* epilogue: records -> [dim][chain] state, counters, save point outputs */
    for (int li = tid; li < nown; li += BLOCK) {
        const long long ch = base + 2 * li + (int)rank;
        const double* rec = sm.rec + (size_t)li * RS;
#pragma unroll
        for (int j = 0; j < DMAX; ++j) {
            const double v = rec[j];
            const long long o = (long long)j * a.st.pitch + ch;
            a.st.X[o] = v;
            if (a.sv.out) a.sv.out[(long long)j * a.sv.out_pitch + ch] = v;
            if (a.sv.sum) save_moments(a.sv, o, v);
        }
        const double lpv = rec[DMAX];
        const unsigned na = sm.nacc[li];
        a.st.lp[ch] = lpv;
        a.st.acc[ch] = (unsigned char)(na >> 31);
        a.st.nacc[ch] = a.st.nacc[ch] + (unsigned long long)(na & 0x7fffffffu);
        if (a.sv.out) a.sv.out[(long long)DMAX * a.sv.out_pitch + ch] = lpv;
        if (a.sv.acc_out) a.sv.acc_out[ch] = (unsigned char)(na >> 31);
    }
    cluster_arrive(); cluster_wait();                        /* shared memory must outlive the other CTA's reads */
}

/* host side: returns true when K2R took the launch (rc = its status) */
template <int DMAX, class T>
bool launch_stretch_res_t(amh_run& r, int nsteps, const StretchArgs& a, const typename T::template Params<DMAX>& tp, int* rc) {
    constexpr int PB = 512;                      /* two plan CTAs per SM: one computes while the other sits at a barrier */
    /* threads of a sweep CTA: 384 = 12 warps x 168 registers.  512 x 128 spills the one-window-ahead loads (8.7e9 vs 1.05e10
     * moves/s on config 3); the level-1 chunks of config 3 need two rounds with 12 as with 16 warps */
    constexpr int BLOCK = 384;
    static const char* res_env = std::getenv("AMH_STRETCH_RES");                       /* A/B / test switch: 0 = K2F, 1 = K2R whenever it fits */
    const long long nw = a.n_walkers;
    if (nw < 64 || nw > 16384 || nsteps >= 65535) return false;
    const long long nens = r.n / nw;
    /* two SMs per ensemble pay when the ensembles are few and large (BASELINE config 3: 64 x 4 096 -> 128 of 148 SMs);
     * many small ensembles are faster one CTA each (K2F): measured in profiles/r2_k2r_vs_k2f_shapes.txt */
    /* ... and when the target is cheap next to the hand-offs: with a dense quadratic form of d >= 12 per move (MvNormal /
     * GaussPrec) the 2 x 384 threads of a cluster lose to K2F's wider CTA (d = 12 / 16, 64 x 2 048 walkers: 5.6 vs 6.1 and
     * 4.0 vs 4.6e9 moves/s, profiles/r2_c3_shape_sweep.txt) */
    constexpr bool dense_target = std::is_same<T, TMvNormal>::value || std::is_same<T, TGaussPrec>::value;
    if (res_env ? std::atoi(res_env) == 0 : !(stretch_res_shape(nw, nens, r.ctx->sm_count) && !(dense_target && DMAX >= 12))) return false;
    const size_t nwl = (size_t)(nw + 1) / 2;
    const size_t rec_b = ((nwl * (DMAX + 1) + 1) & ~(size_t)1) * sizeof(double);
    const size_t fixed_b = rec_b + nwl * (sizeof(unsigned) + sizeof(unsigned short)) + 64;
    const size_t budget = 226 * 1024;
    if (fixed_b + 16 * (DMAX * 8 + 8) > budget) return false;                           /* does not fit: K2F */
    long long fcap = (long long)((budget - fixed_b) / (DMAX * 8 + 8));
    fcap = std::min<long long>(fcap, 4094);
    if (const char* ev = std::getenv("AMH_STRETCH_FWD")) fcap = std::max<long long>(0, std::min<long long>(fcap, std::atoll(ev)));   /* test switch */
    int lcap = kStretchLevels;
    if (const char* ev = std::getenv("AMH_STRETCH_LEVELS")) {                          /* test switch: forces the overflow bucket */
        const int v = std::atoi(ev);
        if (v >= 2 && v <= kStretchLevels) lcap = v;
    }
    int wincap = BLOCK;
    if (const char* ev = std::getenv("AMH_STRETCH_WIN")) {                             /* test switch: many small level-0 windows */
        const int v = std::atoi(ev);
        if (v >= 1 && v <= BLOCK) wincap = v;
    }
    while (2 * ((long long)nwl + wincap - 1) / wincap + 1 > kResMaxWin) wincap = std::min(BLOCK, wincap * 2);
    if (2 * ((long long)nwl + wincap - 1) / wincap + 1 > kResMaxWin) return false;
    static const bool no_ahead = std::getenv("AMH_STRETCH_NO_AHEAD") != nullptr;
    auto body = [&]() -> int {
        StretchPlanR plan;
        plan.nwp = res_nwp(nw);
        const size_t slots = (size_t)nsteps * nens * 2 * plan.nwp;
        const size_t pbytes = (slots * 4 * sizeof(double) + (size_t)nsteps * nens * 2 * kResMeta * sizeof(int) + 255) & ~(size_t)255;
        const size_t gbytes = ((size_t)r.n * res_gs(DMAX) * sizeof(double) + 255) & ~(size_t)255;
        const size_t need = 256 + gbytes + 2 * pbytes;
        if (!r.aux_stream) {
            AMH_CUDA_TRY(cudaStreamCreateWithFlags(&r.aux_stream, cudaStreamNonBlocking));
            for (int b = 0; b < 2; ++b) {
                AMH_CUDA_TRY(cudaEventCreateWithFlags(&r.ev_plan[b], cudaEventDisableTiming));
                AMH_CUDA_TRY(cudaEventCreateWithFlags(&r.ev_sweep[b], cudaEventDisableTiming));
            }
        }
        if (r.plan_layout_nsteps != nsteps) {
            AMH_CUDA_TRY(cudaStreamSynchronize(r.ctx->stream));
            AMH_CUDA_TRY(cudaStreamSynchronize(r.aux_stream));
            r.plan_step0[0] = r.plan_step0[1] = -1;
            r.plan_layout_nsteps = nsteps;
        }
        if (need > r.scratch_bytes) {
            AMH_CUDA_TRY(cudaStreamSynchronize(r.aux_stream));
            r.plan_step0[0] = r.plan_step0[1] = -1;
            dfree(r.ctx, r.scratch);
            r.scratch = nullptr; r.scratch_bytes = 0;
            const int rca = dmalloc(r.ctx, &r.scratch, need);
            if (rca) return rca;
            r.scratch_bytes = need;
        }
        double* Gall = (double*)(((uintptr_t)r.scratch + 255) & ~(uintptr_t)255);
        char* pbase = (char*)Gall + gbytes;
        auto plan_at = [&](int b) {
            StretchPlanR q = plan;
            char* p0 = pbase + (size_t)b * pbytes;
            q.ent = (double*)p0;
            q.meta = (int*)(q.ent + slots * 4);
            return q;
        };
        const size_t smemp = (size_t)nw * 24 + 16;
        auto kp = stretch_plan_res_kernel<PB>;
        if (smemp > 40 * 1024) AMH_CUDA_TRY(cudaFuncSetAttribute(kp, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemp));
        auto enqueue_plan = [&](cudaStream_t st, int b, unsigned long long step0) -> int {
            kp<<<(unsigned)(nsteps * nens), PB, smemp, st>>>(plan_at(b), r.seeds, r.n, (int)nw, r.dim, step0, a.a, lcap, (int)fcap, wincap);
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
                AMH_CUDA_TRY(cudaStreamWaitEvent(r.ctx->stream, r.ev_plan[cur], 0));
            } else {
                cur = 0;
                if (r.plan_step0[0] >= 0) AMH_CUDA_TRY(cudaStreamWaitEvent(r.ctx->stream, r.ev_plan[0], 0));
                const int rcp = enqueue_plan(r.ctx->stream, 0, a.step0);
                if (rcp) return rcp;
            }
        }
        plan = plan_at(cur >= 0 ? cur : 0);
        const size_t smemv = fixed_b + (size_t)fcap * (DMAX * 8 + 8);
        auto kf = stretch_sweep_res_kernel<DMAX, T, BLOCK>;
        if (smemv > 48 * 1024) AMH_CUDA_TRY(cudaFuncSetAttribute(kf, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemv));
        kf<<<(unsigned)(2 * nens), BLOCK, smemv, r.ctx->stream>>>(a, plan, tp, Gall, (int)fcap);
        AMH_CUDA_TRY(cudaGetLastError());
        r.launches += 1;
        r.pending_launches += 1;
        if (cur >= 0) {
            AMH_CUDA_TRY(cudaEventRecord(r.ev_sweep[cur], r.ctx->stream));
            if (!no_ahead) {
                const int o = cur ^ 1;
                AMH_CUDA_TRY(cudaStreamWaitEvent(r.aux_stream, r.ev_sweep[o], 0));
                const int rcp = enqueue_plan(r.aux_stream, o, a.step0 + (unsigned long long)nsteps);
                if (rcp) return rcp;
            }
        }
        return AMH_OK;
    };
    *rc = body();
    return true;
}

}  // namespace amhh
