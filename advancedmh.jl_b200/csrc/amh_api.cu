/* amh_api.cu -- host runtime behind the C ABI of include/amh.h: contexts,
 * target / sampler objects, run state in HBM, the AbstractMCMC sampling
 * schedule and sample / summary read-back.  There is no CPU code path for the
 * MCMC arithmetic in this library: without a CUDA device amh_ctx_create fails.
 */
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <limits>
#include "amh_host.h"

namespace amhh {

static thread_local std::string g_err;

int fail(int code, const std::string& msg) {
    g_err = msg;
    return code;
}
int cuda_fail(cudaError_t e, const char* what) {
    g_err = std::string("CUDA error: ") + cudaGetErrorString(e) + " in " + what;
    return AMH_ERR_CUDA;
}

amhd::ChainState chain_state(amh_run& r) {
    amhd::ChainState st;
    st.X = r.X; st.lp = r.lp; st.lq = r.lq; st.G = r.G;
    st.acc = r.acc; st.nacc = r.nacc; st.seeds = r.seeds;
    st.n = r.n; st.pitch = r.pitch; st.cv = r.cv;
    return st;
}

int default_steps_per_launch(const amh_run& r) {
    switch (r.sampler->d.kind) {
    case AMH_SAMPLER_STRETCH: return stretch_default_steps_per_launch(r);
    case AMH_SAMPLER_RAM: return r.ram_warp ? 16 : 1;   /* K4W keeps the factor in shared memory across fused steps */
    default: return 64;
    }
}
/* amh_run_sample: the steps between two save points (a whole thinning interval) go into ONE launch where the kernel
 * keeps its state on chip across fused steps (MH / MALA: the launch's only HBM traffic is the state at both ends);
 * capped so that a launch stays in the millisecond range.  Stretch plans and K4W tiles are sized per launch length
 * and keep their defaults. */
static int sample_steps_per_launch(const amh_run& r, long long interval) {
    switch (r.sampler->d.kind) {
    case AMH_SAMPLER_STRETCH:
    case AMH_SAMPLER_RAM: return default_steps_per_launch(r);
    default: return (int)std::max<long long>(1, std::min<long long>(interval, 1024));
    }
}

cudaError_t sync_stream(amh_ctx* ctx, cudaStream_t st) {
    if (!ctx->blocking_wait) return cudaStreamSynchronize(st);
    if (!ctx->wait_ev) {
        const cudaError_t e = cudaEventCreateWithFlags(&ctx->wait_ev, cudaEventBlockingSync | cudaEventDisableTiming);
        if (e != cudaSuccess) return e;
    }
    const cudaError_t e = cudaEventRecord(ctx->wait_ev, st);
    if (e != cudaSuccess) return e;
    return cudaEventSynchronize(ctx->wait_ev);
}
void ctx_set_blocking_wait(amh_ctx* ctx, bool on) { if (ctx) ctx->blocking_wait = on; }

int dmalloc(amh_ctx* ctx, void** p, size_t bytes) {
    *p = nullptr;
    if (bytes == 0) return AMH_OK;
    if (ctx->pool) AMH_CUDA_TRY(cudaMallocFromPoolAsync(p, bytes, ctx->pool, ctx->stream));
    else AMH_CUDA_TRY(cudaMallocAsync(p, bytes, ctx->stream));
    return AMH_OK;
}
void dfree(amh_ctx* ctx, void* p) {
    if (p) cudaFreeAsync(p, ctx->stream);
}

template <class T>
static int dev_alloc(amh_ctx* ctx, T** p, size_t count) {
    return dmalloc(ctx, (void**)p, count * sizeof(T));
}

static int upload(amh_ctx* ctx, double** dptr, const std::vector<double>& h) {
    *dptr = nullptr;
    if (h.empty()) return AMH_OK;
    int rc = dmalloc(ctx, (void**)dptr, h.size() * sizeof(double));
    if (rc) return rc;
    AMH_CUDA_TRY(cudaMemcpyAsync(*dptr, h.data(), h.size() * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    AMH_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return AMH_OK;
}

static int get_events(amh_run& r, cudaEvent_t* a, cudaEvent_t* b) {
    if (!r.pool.empty()) {
        *a = r.pool.back().first; *b = r.pool.back().second;
        r.pool.pop_back();
        return AMH_OK;
    }
    AMH_CUDA_TRY(cudaEventCreate(a));
    AMH_CUDA_TRY(cudaEventCreate(b));
    return AMH_OK;
}

static int resolve_events(amh_run& r) {
    if (r.pending.empty()) return AMH_OK;
    AMH_CUDA_TRY(cudaEventSynchronize(r.pending.back().second));
    for (auto& pr : r.pending) {
        float ms = 0;
        AMH_CUDA_TRY(cudaEventElapsedTime(&ms, pr.first, pr.second));
        r.kernel_ms += ms;
        r.pool.push_back(pr);
    }
    r.pending.clear();
    r.timed_launches += r.pending_launches;
    r.pending_launches = 0;
    return AMH_OK;
}

/* enqueue `nsteps` stateful steps; the last launch carries the save epilogue */
static int enqueue_steps(amh_run& r, long long nsteps, bool warmup, int spl, const amhd::SaveArgs* sv_last) {
    if (spl <= 0) spl = default_steps_per_launch(r);
    amhd::SaveArgs none;
    std::memset(&none, 0, sizeof(none));
    AMH_CUDA_TRY(cudaSetDevice(r.ctx->device));
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    int rc = AMH_OK;
    if (r.timing) {
        rc = get_events(r, &e0, &e1);
        if (rc) return rc;
        AMH_CUDA_TRY(cudaEventRecord(e0, r.ctx->stream));
    }
    long long left = nsteps;
    bool first = true;
    while (left > 0 || (first && sv_last)) {
        const int m = (int)std::min<long long>(left, spl);
        const bool last = (left - m) == 0;
        const amhd::SaveArgs& sv = (last && sv_last) ? *sv_last : none;
        switch (r.sampler->d.kind) {
        case AMH_SAMPLER_STATIC:
        case AMH_SAMPLER_RW:
        case AMH_SAMPLER_MIXED: rc = r.sampler->by_components() ? launch_mh_comp(r, m, sv) : launch_mh(r, m, sv); break;
        case AMH_SAMPLER_MALA: rc = launch_mala(r, m, sv); break;
        case AMH_SAMPLER_RAM: rc = r.ram_warp ? launch_ram_warp(r, m, warmup, sv) : launch_ram(r, m, warmup, sv); break;
        case AMH_SAMPLER_STRETCH: rc = launch_stretch(r, m, sv); break;
        default: rc = fail(AMH_ERR_INVALID, "unknown sampler kind");
        }
        if (rc) return rc;
        r.step += m;
        left -= m;
        first = false;
    }
    if (r.timing) {
        AMH_CUDA_TRY(cudaEventRecord(e1, r.ctx->stream));
        r.pending.emplace_back(e0, e1);
        if (r.pending.size() > 2048) {
            rc = resolve_events(r);
            if (rc) return rc;
        }
    } else {
        r.pending_launches = 0;
    }
    return AMH_OK;
}

static void free_run(amh_run* r) {
    if (!r) return;
    cudaSetDevice(r->ctx->device);
    cudaStreamSynchronize(r->ctx->stream);
    if (r->aux_stream) {
        cudaStreamSynchronize(r->aux_stream);
        cudaStreamDestroy(r->aux_stream);
        for (int b = 0; b < 2; ++b) {
            if (r->ev_plan[b]) cudaEventDestroy(r->ev_plan[b]);
            if (r->ev_sweep[b]) cudaEventDestroy(r->ev_sweep[b]);
        }
    }
    void* ptrs[] = {r->X, r->X2, r->lp, r->lp2, r->lq, r->G, r->S, r->S2, r->logalpha, r->eta, r->acc, r->failed,
                    r->sflag, r->nacc, r->seeds, r->sum, r->sumsq, r->scratch, r->scratch2};
    for (void* p : ptrs) dfree(r->ctx, p);
    mala_tensor_release(*r);
    for (auto& pr : r->pending) { cudaEventDestroy(pr.first); cudaEventDestroy(pr.second); }
    for (auto& pr : r->pool) { cudaEventDestroy(pr.first); cudaEventDestroy(pr.second); }
    delete r;
}

}  // namespace amhh

using namespace amhh;

extern "C" {

int32_t amh_version(int32_t* major, int32_t* minor) {
    if (major) *major = AMH_VERSION_MAJOR;
    if (minor) *minor = AMH_VERSION_MINOR;
    return AMH_OK;
}
const char* amh_last_error(void) { return g_err.c_str(); }
int32_t amh_contract_version(void) { return AMH_CONTRACT_VERSION; }

int32_t amh_ctx_create(int32_t device, amh_ctx** out) {
    if (!out) return fail(AMH_ERR_INVALID, "out is NULL");
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return fail(AMH_ERR_CUDA, std::string("no CUDA device available (this library has no CPU fallback): ") +
                                      cudaGetErrorString(e));
    if (device < 0 || device >= count) return fail(AMH_ERR_INVALID, "device index out of range");
    AMH_CUDA_TRY(cudaSetDevice(device));
    amh_ctx* c = new amh_ctx();
    c->device = device;
    AMH_CUDA_TRY(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    AMH_CUDA_TRY(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
    AMH_CUDA_TRY(cudaDeviceGetAttribute(&c->sm_count, cudaDevAttrMultiProcessorCount, device));
    {   /* a stream-ordered pool PER CONTEXT: freed blocks stay cached (run handles are created and destroyed per
         * `sample` call), and a block is only ever reused on the stream that freed it -- with the device's default
         * pool the driver orders a context's stream behind another context's to recycle its memory, which serialises
         * contexts that are meant to run concurrently (MCMCB200(streams=k)) */
        cudaMemPoolProps props;
        std::memset(&props, 0, sizeof(props));
        props.allocType = cudaMemAllocationTypePinned;
        props.handleTypes = cudaMemHandleTypeNone;
        props.location.type = cudaMemLocationTypeDevice;
        props.location.id = device;
        AMH_CUDA_TRY(cudaMemPoolCreate(&c->pool, &props));
        unsigned long long keep = ~0ull;
        AMH_CUDA_TRY(cudaMemPoolSetAttribute(c->pool, cudaMemPoolAttrReleaseThreshold, &keep));
    }
    *out = c;
    return AMH_OK;
}
int32_t amh_ctx_destroy(amh_ctx* c) {
    if (!c) return AMH_OK;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    if (c->wait_ev) cudaEventDestroy(c->wait_ev);
    cudaStreamDestroy(c->stream);
    cudaStreamDestroy(c->copy_stream);
    if (c->pool) cudaMemPoolDestroy(c->pool);
    delete c;
    return AMH_OK;
}
int32_t amh_ctx_sync(amh_ctx* c) {
    if (!c) return fail(AMH_ERR_INVALID, "ctx is NULL");
    AMH_CUDA_TRY(cudaSetDevice(c->device));
    AMH_CUDA_TRY(cudaStreamSynchronize(c->stream));
    return AMH_OK;
}

int32_t amh_target_create(amh_ctx* ctx, int32_t kind, int32_t dim, const double* blob, int64_t nblob,
                          amh_target** out) {
    return amhh::target_create_impl(ctx, kind, dim, blob, nblob, out, true);
}
}  /* extern "C" */

/* upload_data == false: device memory is allocated but left for a broadcast to fill (amh_job.cu) */
int amhh::target_create_impl(amh_ctx* ctx, int32_t kind, int32_t dim, const double* blob, int64_t nblob,
                             amh_target** out, bool upload_data) {
    if (!ctx || !out) return fail(AMH_ERR_INVALID, "ctx/out is NULL");
    if (dim < 1) return fail(AMH_ERR_INVALID, "dim must be >= 1");
    if (nblob < 0 || (nblob > 0 && !blob)) return fail(AMH_ERR_INVALID, "blob is NULL");
    const long long d = dim;
    bool ok = true;
    long long ndata = 0;
    double inv2tau2 = 0, invtau2 = 0;
    switch (kind) {
    case AMH_TARGET_IID_NORMAL: ok = (dim == 2 && nblob >= 1); ndata = nblob; break;
    case AMH_TARGET_MVNORMAL: ok = (nblob == 1 + d + d * (d + 1) / 2); break;
    case AMH_TARGET_ROSENBROCK: ok = (nblob == 3 && dim >= 2); break;
    case AMH_TARGET_GAUSS_PREC: ok = (nblob == d * d); break;
    case AMH_TARGET_NIG_TOY:
    case AMH_TARGET_NIG_TOY_LOG: ok = (dim == 2 && nblob >= 3); ndata = nblob - 3; break;
    case AMH_TARGET_LOGISTIC:
        ok = (nblob >= 1 + d + 1) && ((nblob - 1) % (d + 1) == 0);
        if (ok) {
            ndata = (nblob - 1) / (d + 1);
            const double tau = blob[0];
            inv2tau2 = 1.0 / (2.0 * tau * tau);
            invtau2 = 1.0 / (tau * tau);
        }
        break;
    default: ok = false;
    }
    if (!ok) return fail(AMH_ERR_INVALID, "target kind/dim/blob size mismatch");
    AMH_CUDA_TRY(cudaSetDevice(ctx->device));
    amh_target* t = new amh_target();
    t->ctx = ctx; t->kind = kind; t->dim = dim; t->ndata = ndata;
    t->inv2tau2 = inv2tau2; t->invtau2 = invtau2;
    t->blob.assign(blob, blob + nblob);
    const int rc = upload_data ? upload(ctx, &t->dblob, t->blob) : dmalloc(ctx, (void**)&t->dblob, sizeof(double) * (size_t)nblob);
    if (rc) { delete t; return rc; }
    *out = t;
    return AMH_OK;
}
int amhh::target_create_empty(amh_ctx* ctx, int32_t kind, int32_t dim, const double* blob, int64_t nblob, amh_target** out) {
    return target_create_impl(ctx, kind, dim, blob, nblob, out, false);
}

extern "C" {
int32_t amh_target_create_source(amh_ctx* ctx, int32_t dim, const char* source, int32_t has_gradient,
                                 const double* data, int64_t ndata, amh_target** out) {
    if (!ctx || !out) return fail(AMH_ERR_INVALID, "ctx/out is NULL");
    if (!source || !*source) return fail(AMH_ERR_INVALID, "source is NULL or empty");
    if (dim < 1) return fail(AMH_ERR_INVALID, "dim must be >= 1");
    if (dim > amhd::kGenericCap) return fail(AMH_ERR_UNSUPPORTED, "user-supplied targets support dim <= 128");
    if (ndata < 0 || (ndata > 0 && !data)) return fail(AMH_ERR_INVALID, "data is NULL");
    AMH_CUDA_TRY(cudaSetDevice(ctx->device));
    amh_target* t = new amh_target();
    t->ctx = ctx; t->kind = AMH_TARGET_USER; t->dim = dim; t->ndata = ndata;
    t->user_grad = has_gradient != 0;
    if (ndata > 0) t->blob.assign(data, data + ndata);
    else t->blob.assign(1, 0.0);                      /* the kernels always get a valid pointer */
    int rc = upload(ctx, &t->dblob, t->blob);
    if (!rc) rc = rtc_build(*t, source, t->user_grad);
    if (rc) { amh_target_destroy(t); return rc; }
    *out = t;
    return AMH_OK;
}
int32_t amh_target_destroy(amh_target* t) {
    if (!t) return AMH_OK;
    cudaSetDevice(t->ctx->device);
    if (t->rtc) {
        cudaStreamSynchronize(t->ctx->stream);        /* no kernel of the module may still be running */
        rtc_destroy(*t);
    }
    dfree(t->ctx, t->dblob);
    delete t;
    return AMH_OK;
}

/* parameter validation of one univariate law (Distributions.jl constructors throw DomainError) */
static const char* check_component(const amh_component& q) {
    switch (q.family) {
    case AMH_FAM_NORMAL:
    case AMH_FAM_LOGNORMAL:   return (q.p1 > 0.0) ? nullptr : "Normal / LogNormal need sigma > 0";
    case AMH_FAM_INVGAMMA:
    case AMH_FAM_GAMMA:       return (q.p0 > 0.0 && q.p1 > 0.0) ? nullptr : "Gamma / InverseGamma need shape > 0 and scale > 0";
    case AMH_FAM_UNIFORM:     return (q.p1 > q.p0) ? nullptr : "Uniform needs a < b";
    case AMH_FAM_EXPONENTIAL: return (q.p0 > 0.0) ? nullptr : "Exponential needs scale > 0";
    }
    return "unknown distribution family";
}

int32_t amh_sampler_create(amh_ctx* ctx, const amh_sampler_desc* desc, amh_sampler** out) {
    if (!ctx || !desc || !out) return fail(AMH_ERR_INVALID, "ctx/desc/out is NULL");
    const int d = desc->dim;
    if (d < 1) return fail(AMH_ERR_INVALID, "dim must be >= 1");
    amh_sampler* s = new amh_sampler();
    s->ctx = ctx;
    s->d = *desc;
    const long long nt = (long long)d * (d + 1) / 2;
    auto bad = [&](const char* m) { delete s; return fail(AMH_ERR_INVALID, m); };
    switch (desc->kind) {
    case AMH_SAMPLER_STATIC:
    case AMH_SAMPLER_RW:
    case AMH_SAMPLER_MIXED:
    case AMH_SAMPLER_STRETCH: {
        if (desc->kind == AMH_SAMPLER_STRETCH) {
            if (desc->n_walkers < 2) return bad("Ensemble needs n_walkers >= 2");
            if (!(desc->stretch_a > 1.0)) return bad("stretch_length must be > 1");
        }
        if (desc->kind == AMH_SAMPLER_MIXED || desc->cov_kind == AMH_COV_COMPONENTS) {
            if (!desc->components) return bad("components is NULL");
            if (d > amhd::kGenericCap) return bad("component proposals support dim <= 128");
            s->comps.assign(desc->components, desc->components + d);
            for (const amh_component& q : s->comps) {
                const char* m = check_component(q);
                if (m) return bad(m);
            }
            break;
        }
        const bool need_cov = desc->kind != AMH_SAMPLER_STRETCH || desc->scale != nullptr;
        if (need_cov) {
            if (!desc->scale) return bad("proposal scale is NULL");
            const long long ns = desc->cov_kind == AMH_COV_FULL ? nt : desc->cov_kind == AMH_COV_DIAG ? d
                                 : desc->cov_kind == AMH_COV_SCALAR ? 1 : -1;
            if (ns < 0) return bad("unknown cov_kind");
            s->scale.assign(desc->scale, desc->scale + ns);
            for (int i = 0; i < d; ++i) {
                const double dg = desc->cov_kind == AMH_COV_FULL ? s->scale[tri_h(i, i)]
                                  : desc->cov_kind == AMH_COV_DIAG ? s->scale[i] : s->scale[0];
                if (!(dg > 0.0)) return bad("proposal scale must have a positive diagonal");
            }
        }
        if (desc->mean) { s->mean.assign(desc->mean, desc->mean + d); s->has_mean = true; }
        break;
    }
    case AMH_SAMPLER_MALA:
        if (!(desc->mala_sigma2 > 0.0)) return bad("MALA sigma2 must be > 0");
        s->mala_sigma = std::sqrt(desc->mala_sigma2);
        break;
    case AMH_SAMPLER_RAM:
        if (desc->ram_S0) {
            s->S0.resize(nt);
            for (int i = 0; i < d; ++i)
                for (int j = 0; j <= i; ++j) s->S0[tri_h(i, j)] = desc->ram_S0[(long long)i * d + j];
        }
        break;
    default:
        return bad("unknown sampler kind");
    }
    s->d.mean = nullptr; s->d.scale = nullptr; s->d.ram_S0 = nullptr; s->d.components = nullptr;
    if (s->d.contract == 0) {
        const char* ev = std::getenv("AMH_CONTRACT");
        s->d.contract = ev ? std::atoi(ev) : AMH_CONTRACT_VERSION;
    }
    if (s->d.contract != AMH_CONTRACT_V1 && s->d.contract != AMH_CONTRACT_V2) return bad("unknown contract version");
    cudaSetDevice(ctx->device);
    int rc = upload(ctx, &s->dmean, s->mean);
    if (!rc && !s->comps.empty()) {
        rc = dmalloc(ctx, (void**)&s->dcomps, s->comps.size() * sizeof(amh_component));
        if (!rc) {
            cudaError_t e = cudaMemcpyAsync(s->dcomps, s->comps.data(), s->comps.size() * sizeof(amh_component),
                                            cudaMemcpyHostToDevice, ctx->stream);
            if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
            if (e != cudaSuccess) rc = cuda_fail(e, "copy components");
        }
    }
    if (!rc) rc = upload(ctx, &s->dscale, s->scale);
    if (!rc) rc = upload(ctx, &s->dS0, s->S0);
    if (rc) { amh_sampler_destroy(s); return rc; }
    *out = s;
    return AMH_OK;
}
int32_t amh_sampler_destroy(amh_sampler* s) {
    if (!s) return AMH_OK;
    cudaSetDevice(s->ctx->device);
    dfree(s->ctx, s->dmean);
    dfree(s->ctx, s->dscale);
    dfree(s->ctx, s->dS0);
    dfree(s->ctx, s->dcomps);
    delete s;
    return AMH_OK;
}

int32_t amh_run_create(amh_ctx* ctx, amh_target* target, amh_sampler* sampler, int64_t n, int64_t off,
                       const uint64_t* seeds, const double* init, int64_t init_ld, amh_run** out) {
    if (!ctx || !target || !sampler || !out || !seeds) return fail(AMH_ERR_INVALID, "NULL argument");
    if (target->dim != sampler->d.dim) return fail(AMH_ERR_INVALID, "target and sampler dimensions differ");
    if (target->ctx != ctx || sampler->ctx != ctx) return fail(AMH_ERR_INVALID, "target/sampler belong to another ctx");
    if (n < 1) return fail(AMH_ERR_INVALID, "nchains_local must be >= 1");
    if (init_ld == 0) init_ld = n;
    if (init && init_ld < n) return fail(AMH_ERR_INVALID, "init_ld must be >= nchains_local");
    const int kind = sampler->d.kind;
    const int d = target->dim;
    if (d > amhd::kGenericCap) return fail(AMH_ERR_UNSUPPORTED, "device samplers support dim <= 128");
    long long nseeds = n;
    if (kind == AMH_SAMPLER_STRETCH) {
        if (n % sampler->d.n_walkers) return fail(AMH_ERR_INVALID, "nchains_local must be a multiple of n_walkers");
        nseeds = n / sampler->d.n_walkers;
    }
    if (kind == AMH_SAMPLER_MALA) {
        /* propose(::MALA) = error("please specify initial parameters")  (MALA.jl:37) */
        if (!init) return fail(AMH_ERR_STATE, "please specify initial parameters");
        /* check_capabilities (MALA.jl:42-52) */
        if (!target->has_grad()) return fail(AMH_ERR_INVALID, "The gradient of the log density function is not defined");
    }
    if (kind == AMH_SAMPLER_STRETCH && !init && sampler->scale.empty() && !sampler->by_components())
        return fail(AMH_ERR_INVALID, "stretch move without init needs an initial-draw proposal");
    if (sampler->d.precision != AMH_PRECISION_FP64 &&
        !(sampler->d.precision == AMH_PRECISION_BF16X2 && kind == AMH_SAMPLER_MALA && target->kind == AMH_TARGET_LOGISTIC && d == 128 &&
          target->ndata >= 64))
        return fail(AMH_ERR_UNSUPPORTED, "precision bf16x2 (split-bf16 tensor-core path) exists for MALA on the logistic target with dim = 128 "
                                         "and >= 64 rows; every other combination runs in fp64");
    AMH_CUDA_TRY(cudaSetDevice(ctx->device));
    amh_run* r = new amh_run();
    r->ctx = ctx; r->target = target; r->sampler = sampler;
    r->n = n; r->off = off; r->dim = d; r->nseeds = nseeds;
    r->cv = sampler->d.contract;
    r->pitch = (n + 31) / 32 * 32;          /* rows start 256-byte aligned; a warp's 32 chains never straddle a row end */
    if (const char* ev = std::getenv("AMH_PITCH_PAD")) r->pitch += 32ll * std::atoll(ev);      /* experiment switch */
    {
        const char* ev = std::getenv("AMH_MH_PATH");     /* developer switch for A/B measurements: "dfma" forces K1 */
        if (ev && std::strcmp(ev, "dfma") == 0) r->mh_path = 1;
        if (ev && std::strcmp(ev, "tc32") == 0) r->mh_path = 2;
    }
    const size_t np = (size_t)r->pitch;
    const size_t nt = (size_t)d * (d + 1) / 2;
    int rc = AMH_OK;
    auto chk = [&](int c) { if (!rc) rc = c; };
    r->x_rows = d <= 64 ? (d + 7) & ~7 : (d + 15) & ~15;   /* padding rows for the padded tensor-core MH kernels (amh_launch_mh_tcp.cu) */
    if (target->kind == AMH_TARGET_LOGISTIC && d <= 128)     /* ... and of the tiled logistic kernels: features padded to 32 / 64 / 128 */
        r->x_rows = d <= 32 ? 32 : d <= 64 ? 64 : 128;
    chk(dev_alloc(ctx, &r->X, (size_t)r->x_rows * np));
    chk(dev_alloc(ctx, &r->lp, np));
    chk(dev_alloc(ctx, &r->lq, np));
    chk(dev_alloc(ctx, &r->acc, np));
    chk(dev_alloc(ctx, &r->failed, np));
    chk(dev_alloc(ctx, &r->nacc, np));
    chk(dev_alloc(ctx, &r->seeds, (size_t)nseeds));
    chk(dev_alloc(ctx, &r->sum, (size_t)d * np));
    chk(dev_alloc(ctx, &r->sumsq, (size_t)d * np));
    if (kind == AMH_SAMPLER_MALA) chk(dev_alloc(ctx, &r->G, (size_t)r->x_rows * np));
    if (kind == AMH_SAMPLER_STRETCH) {
        chk(dev_alloc(ctx, &r->X2, (size_t)d * np));
        chk(dev_alloc(ctx, &r->lp2, np));
    }
    if (kind == AMH_SAMPLER_RAM) {
        const char* rp = std::getenv("AMH_RAM_PATH");          /* A/B switch: thread forces K4 */
        r->ram_warp = ram_warp_eligible(*r) && !(rp && std::strcmp(rp, "thread") == 0);
    }
    if (kind == AMH_SAMPLER_RAM && r->ram_warp) {
        chk(dev_alloc(ctx, &r->S, ((nt + 1) & ~(size_t)1) * (size_t)n));   /* [chain][column-packed, padded to 16 B] */

        chk(dev_alloc(ctx, &r->logalpha, np));
        chk(dev_alloc(ctx, &r->eta, np));
    } else if (kind == AMH_SAMPLER_RAM) {
        chk(dev_alloc(ctx, &r->S, nt * np));
        chk(dev_alloc(ctx, &r->logalpha, np));
        chk(dev_alloc(ctx, &r->eta, np));
        /* second factor buffer + per-chain selector: the non-mutating lowrankupdate/downdate (RAM :167,:170) */
        chk(dev_alloc(ctx, &r->S2, nt * np));
        chk(dev_alloc(ctx, &r->sflag, np));
    }
    if (rc) { free_run(r); return rc; }
    cudaStream_t st = ctx->stream;
    auto cu = [&](cudaError_t e, const char* w) { if (!rc && e != cudaSuccess) rc = cuda_fail(e, w); };
    cu(cudaMemsetAsync(r->X, 0, sizeof(double) * r->x_rows * np, st), "memset X");
    if (r->G) cu(cudaMemsetAsync(r->G, 0, sizeof(double) * r->x_rows * np, st), "memset G");      /* the padding rows must be 0 */
    cu(cudaMemsetAsync(r->lp, 0, sizeof(double) * np, st), "memset lp");
    cu(cudaMemsetAsync(r->lq, 0, sizeof(double) * np, st), "memset lq");
    cu(cudaMemsetAsync(r->acc, 0, np, st), "memset acc");
    cu(cudaMemsetAsync(r->failed, 0, np, st), "memset failed");
    cu(cudaMemsetAsync(r->nacc, 0, sizeof(unsigned long long) * np, st), "memset nacc");
    cu(cudaMemsetAsync(r->sum, 0, sizeof(double) * d * np, st), "memset sum");
    cu(cudaMemsetAsync(r->sumsq, 0, sizeof(double) * d * np, st), "memset sumsq");
    if (r->logalpha) cu(cudaMemsetAsync(r->logalpha, 0, sizeof(double) * np, st), "memset logalpha");
    if (r->eta) cu(cudaMemsetAsync(r->eta, 0, sizeof(double) * np, st), "memset eta");
    if (r->sflag) cu(cudaMemsetAsync(r->sflag, 0, np, st), "memset sflag");
    if (r->S2 && !r->ram_warp) cu(cudaMemsetAsync(r->S2, 0, sizeof(double) * nt * np, st), "memset S2");
    cu(cudaMemcpyAsync(r->seeds, seeds, sizeof(uint64_t) * nseeds, cudaMemcpyHostToDevice, st), "copy seeds");
    int mode;
    if (init) {
        cu(cudaMemcpy2DAsync(r->X, sizeof(double) * np, init, sizeof(double) * init_ld, sizeof(double) * n, d,
                             cudaMemcpyHostToDevice, st), "copy init");
        mode = 0;
    } else {
        mode = kind == AMH_SAMPLER_RAM ? 2 : kind == AMH_SAMPLER_STRETCH ? 3 : 1;
    }
    if (!rc && r->ram_warp) {
        double* Ssave = r->S;
        r->S = nullptr;                  /* the init kernel writes the thread-kernel layout; K4W has its own */
        rc = launch_init(*r, mode);
        r->S = Ssave;
        if (!rc) rc = ramw_init_S(*r);
    } else if (!rc) rc = launch_init(*r, mode);
    cu(amhh::sync_stream(ctx, st), "init sync");      /* host buffers are only read during the call */
    if (rc) { free_run(r); return rc; }
    *out = r;
    return AMH_OK;
}
int32_t amh_run_destroy(amh_run* r) {
    free_run(r);
    return AMH_OK;
}

int32_t amh_run_steps(amh_run* run, int64_t nsteps, int32_t warmup, int32_t steps_per_launch) {
    if (!run) return fail(AMH_ERR_INVALID, "run is NULL");
    if (nsteps < 0) return fail(AMH_ERR_INVALID, "nsteps must be >= 0");
    if (nsteps == 0) return AMH_OK;
    return enqueue_steps(*run, nsteps, warmup != 0, steps_per_launch, nullptr);
}
int32_t amh_run_sync(amh_run* run) {
    if (!run) return fail(AMH_ERR_INVALID, "run is NULL");
    AMH_CUDA_TRY(cudaSetDevice(run->ctx->device));
    AMH_CUDA_TRY(amhh::sync_stream(run->ctx, run->ctx->stream));
    return AMH_OK;
}

int32_t amh_run_sample(amh_run* run, int64_t N, int64_t discard_initial, int64_t thinning, int64_t num_warmup,
                       double* out, uint8_t* accepted_out, amh_summary* summary) {
    if (!run) return fail(AMH_ERR_INVALID, "run is NULL");
    return amh_run_sample_ld(run, N, discard_initial, thinning, num_warmup, out, run->n, accepted_out, run->n, summary);
}

int32_t amh_run_sample_ld(amh_run* run, int64_t N, int64_t discard_initial, int64_t thinning, int64_t num_warmup,
                          double* out, int64_t out_ld, uint8_t* accepted_out, int64_t acc_ld, amh_summary* summary) {
    if (!run) return fail(AMH_ERR_INVALID, "run is NULL");
    if (N < 1 || thinning < 1 || discard_initial < 0 || num_warmup < 0)
        return fail(AMH_ERR_INVALID, "need N >= 1, thinning >= 1, discard_initial >= 0, num_warmup >= 0");
    if ((out && out_ld < run->n) || (accepted_out && acc_ld < run->n))
        return fail(AMH_ERR_INVALID, "out_ld / acc_ld must be >= nchains_local");
    amh_run& r = *run;
    const long long n = r.n, np = r.pitch;
    const int d = r.dim;
    cudaStream_t st = r.ctx->stream;
    static const bool trace = std::getenv("AMH_TRACE") != nullptr;
    const auto tr0 = std::chrono::steady_clock::now();
    auto lap = [&](const char* what) {
        if (trace) std::fprintf(stderr, "[amh] dev %d run_sample %-22s %8.3f ms\n", r.ctx->device, what,
                                std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - tr0).count());
    };
    if (trace) r.timing = true;                  /* the trace also reports the device time of the stepping kernels */
    AMH_CUDA_TRY(cudaSetDevice(r.ctx->device));
    AMH_CUDA_TRY(cudaMemsetAsync(r.sum, 0, sizeof(double) * d * np, st));
    AMH_CUDA_TRY(cudaMemsetAsync(r.sumsq, 0, sizeof(double) * d * np, st));
    r.nsaved = 0;
    /* device sample ring: two buffers of `chunk` slabs of [(d+1)][pitch] doubles (+ accepted flags).  While the
     * stepping kernels fill one buffer, the copy stream drains the other into the caller's (ideally pinned) array. */
    const size_t slab = (size_t)(d + 1) * np;
    long long chunk = 0;
    int nbuf = 0;
    double* dsamp[2] = {nullptr, nullptr};
    unsigned char* dacc[2] = {nullptr, nullptr};
    cudaEvent_t filled[2] = {nullptr, nullptr}, copied[2] = {nullptr, nullptr};
    cudaStream_t cs = r.ctx->copy_stream;
    int rc = AMH_OK;
    if (out || accepted_out) {
        chunk = std::max<long long>(1, std::min<long long>(N, (long long)((128ull << 20) / (slab * sizeof(double)))));
        if (N >= 2) chunk = std::min<long long>(chunk, (N + 1) / 2);   /* at least two chunks: copies overlap stepping */
        nbuf = (N > chunk) ? 2 : 1;
        for (int b = 0; b < nbuf && !rc; ++b) {
            if (out) rc = dmalloc(r.ctx, (void**)&dsamp[b], sizeof(double) * slab * chunk);
            if (!rc && accepted_out) rc = dmalloc(r.ctx, (void**)&dacc[b], (size_t)np * chunk);
            if (!rc && cudaEventCreateWithFlags(&filled[b], cudaEventDisableTiming) != cudaSuccess) rc = fail(AMH_ERR_CUDA, "event");
            if (!rc && cudaEventCreateWithFlags(&copied[b], cudaEventDisableTiming) != cudaSuccess) rc = fail(AMH_ERR_CUDA, "event");
        }
    }
    auto cleanup = [&]() {
        amhh::sync_stream(r.ctx, cs);
        for (int b = 0; b < 2; ++b) {
            dfree(r.ctx, dsamp[b]);
            dfree(r.ctx, dacc[b]);
            if (filled[b]) cudaEventDestroy(filled[b]);
            if (copied[b]) cudaEventDestroy(copied[b]);
        }
    };
    auto drain = [&](int b, long long base, long long count) -> int {
        /* copy stream: wait until the buffer is filled, copy it out, mark it reusable */
        AMH_CUDA_TRY(cudaEventRecord(filled[b], st));
        AMH_CUDA_TRY(cudaStreamWaitEvent(cs, filled[b], 0));
        if (out)
            AMH_CUDA_TRY(cudaMemcpy2DAsync(out + (size_t)base * (d + 1) * out_ld, sizeof(double) * out_ld, dsamp[b], sizeof(double) * np,
                                           sizeof(double) * n, (size_t)count * (d + 1), cudaMemcpyDeviceToHost, cs));
        if (accepted_out)
            AMH_CUDA_TRY(cudaMemcpy2DAsync(accepted_out + (size_t)base * acc_ld, (size_t)acc_ld, dacc[b], (size_t)np, (size_t)n,
                                           (size_t)count, cudaMemcpyDeviceToHost, cs));
        AMH_CUDA_TRY(cudaEventRecord(copied[b], cs));
        return AMH_OK;
    };
    lap("setup done");
    cudaEvent_t tr_ev = nullptr;
    if (trace && cudaEventCreateWithFlags(&tr_ev, cudaEventDisableTiming) == cudaSuccess) cudaEventRecord(tr_ev, st);
    for (long long i = 0; i < N && !rc; ++i) {
        long long k = (i == 0) ? discard_initial : thinning;
        const int b = chunk ? (int)((i / chunk) % nbuf) : 0;
        const long long slot = chunk ? i % chunk : 0;
        if (chunk && slot == 0 && i >= (long long)nbuf * chunk) {
            cudaError_t e = cudaStreamWaitEvent(st, copied[b], 0);       /* the buffer has been drained */
            if (e != cudaSuccess) { rc = cuda_fail(e, "wait copied"); break; }
        }
        amhd::SaveArgs sv;
        std::memset(&sv, 0, sizeof(sv));
        sv.out = dsamp[b] ? dsamp[b] + (size_t)slot * slab : nullptr;
        sv.out_pitch = np;
        sv.acc_out = dacc[b] ? dacc[b] + (size_t)slot * np : nullptr;
        sv.sum = r.sum;
        sv.sumsq = r.sumsq;
        sv.inv_n = 1.0 / (double)(r.nsaved + 1);
        /* stateful step s (1-based, cumulative) is step_warmup iff s <= num_warmup */
        bool done = false;
        while (!done && !rc) {
            const bool wu = k > 0 && r.step < num_warmup;
            const long long m = wu ? std::min<long long>(k, num_warmup - r.step) : k;
            rc = enqueue_steps(r, m, wu, sample_steps_per_launch(r, m), (m == k) ? &sv : nullptr);
            k -= m;
            done = (k == 0);
        }
        r.nsaved += 1;
        if (chunk && !rc && (slot == chunk - 1 || i == N - 1)) rc = drain(b, i - slot, slot + 1);
    }
    lap("all enqueued");
    if (trace) {
        if (tr_ev) { cudaEventSynchronize(tr_ev); cudaEventDestroy(tr_ev); lap("stream reached run_sample (run_create's copies and init kernel done)"); }
        if (std::getenv("AMH_TRACE_POLL")) { while (cudaStreamQuery(st) == cudaErrorNotReady) {} }   /* spin instead of the driver's wait */
        cudaStreamSynchronize(st);
        lap("step stream idle");
        double kms = 0;
        int64_t kl = 0;
        amh_run_kernel_time_ms(run, 0, &kms, &kl);
        std::fprintf(stderr, "[amh] dev %d run_sample stepping kernels: %.3f ms in %lld launches (CUDA events)\n", r.ctx->device, kms, (long long)kl);
    }
    if (!rc) {
        cudaError_t e = amhh::sync_stream(r.ctx, cs);
        if (e != cudaSuccess) rc = cuda_fail(e, "copy stream sync");
    }
    lap("copies done");
    cleanup();
    lap("cleanup done");
    if (rc) return rc;
    AMH_CUDA_TRY(amhh::sync_stream(r.ctx, st));
    if (summary) {
        std::vector<double> hs((size_t)d * np), hq((size_t)d * np);
        std::vector<unsigned long long> ha((size_t)np);
        AMH_CUDA_TRY(cudaMemcpy(hs.data(), r.sum, sizeof(double) * d * np, cudaMemcpyDeviceToHost));
        AMH_CUDA_TRY(cudaMemcpy(hq.data(), r.sumsq, sizeof(double) * d * np, cudaMemcpyDeviceToHost));
        AMH_CUDA_TRY(cudaMemcpy(ha.data(), r.nacc, sizeof(unsigned long long) * np, cudaMemcpyDeviceToHost));
        summary->n_saved = r.nsaved;
        summary->n_steps = r.step;
        double na = 0;
        for (long long c = 0; c < n; ++c) na += (double)ha[c];
        summary->accept_rate = r.step > 0 ? na / ((double)n * (double)r.step) : 0.0;
        /* per-chain Welford (mean, M2) -> pooled moments (every chain holds n_saved samples): the pooled mean is the
         * mean of the chain means, the pooled M2 is sum M2_c + n_saved * sum (mean_c - mean)^2  (Chan et al.) */
        for (int i = 0; i < d; ++i) {
            double s1 = 0;
            for (long long c = 0; c < n; ++c) {
                s1 += hs[(size_t)i * np + c];
                if (summary->chain_mean) summary->chain_mean[(size_t)i * n + c] = hs[(size_t)i * np + c];
            }
            const double m = s1 / (double)n;
            double m2 = 0, dev = 0;
            for (long long c = 0; c < n; ++c) {
                const double dl = hs[(size_t)i * np + c] - m;
                m2 += hq[(size_t)i * np + c];
                dev = std::fma(dl, dl, dev);
            }
            if (summary->mean) summary->mean[i] = m;
            if (summary->var) summary->var[i] = (m2 + (double)r.nsaved * dev) / ((double)n * (double)r.nsaved);
        }
    }
    return AMH_OK;
}

int32_t amh_run_get_state(amh_run* run, double* x, double* lp, double* grad, double* S, uint8_t* accepted,
                          int64_t* naccept, int64_t* step_counter) {
    if (!run) return fail(AMH_ERR_INVALID, "run is NULL");
    return amh_run_get_state_ld(run, run->n, x, lp, grad, S, accepted, naccept, step_counter);
}

int32_t amh_run_get_state_ld(amh_run* run, int64_t ld, double* x, double* lp, double* grad, double* S, uint8_t* accepted,
                             int64_t* naccept, int64_t* step_counter) {
    if (!run) return fail(AMH_ERR_INVALID, "run is NULL");
    amh_run& r = *run;
    const long long n = r.n, np = r.pitch;
    const int d = r.dim;
    if (ld < n) return fail(AMH_ERR_INVALID, "ld must be >= nchains_local");
    AMH_CUDA_TRY(cudaSetDevice(r.ctx->device));
    AMH_CUDA_TRY(cudaStreamSynchronize(r.ctx->stream));
    if (x) AMH_CUDA_TRY(cudaMemcpy2D(x, sizeof(double) * ld, r.X, sizeof(double) * np, sizeof(double) * n, d, cudaMemcpyDeviceToHost));
    if (lp) AMH_CUDA_TRY(cudaMemcpy(lp, r.lp, sizeof(double) * n, cudaMemcpyDeviceToHost));
    if (grad) {
        if (!r.G) return fail(AMH_ERR_INVALID, "sampler keeps no gradient");
        AMH_CUDA_TRY(cudaMemcpy2D(grad, sizeof(double) * ld, r.G, sizeof(double) * np, sizeof(double) * n, d, cudaMemcpyDeviceToHost));
    }
    if (S) {
        if (!r.S) return fail(AMH_ERR_INVALID, "sampler keeps no Cholesky factor");
        const size_t nt = (size_t)d * (d + 1) / 2;
        double* tmp = nullptr;
        int rc = dmalloc(r.ctx, (void**)&tmp, sizeof(double) * nt * np);
        if (rc) return rc;
        rc = r.ram_warp ? ramw_export_S(r, tmp) : ram_gather_S(r, tmp);
        if (!rc) {
            cudaError_t e = cudaStreamSynchronize(r.ctx->stream);
            if (e == cudaSuccess)
                e = cudaMemcpy2D(S, sizeof(double) * ld, tmp, sizeof(double) * np, sizeof(double) * n, nt, cudaMemcpyDeviceToHost);
            if (e != cudaSuccess) rc = cuda_fail(e, "copy S");
        }
        dfree(r.ctx, tmp);
        if (rc) return rc;
    }
    if (accepted) AMH_CUDA_TRY(cudaMemcpy(accepted, r.acc, (size_t)n, cudaMemcpyDeviceToHost));
    if (naccept) AMH_CUDA_TRY(cudaMemcpy(naccept, r.nacc, sizeof(int64_t) * n, cudaMemcpyDeviceToHost));
    if (step_counter) *step_counter = r.step;
    return AMH_OK;
}

int32_t amh_run_set_state(amh_run* run, const double* x, const double* lp, const double* grad, const double* S,
                          const uint8_t* accepted, const int64_t* naccept, int64_t step_counter) {
    if (!run) return fail(AMH_ERR_INVALID, "run is NULL");
    return amh_run_set_state_ld(run, run->n, x, lp, grad, S, accepted, naccept, step_counter);
}

int32_t amh_run_set_state_ld(amh_run* run, int64_t ld, const double* x, const double* lp, const double* grad, const double* S,
                             const uint8_t* accepted, const int64_t* naccept, int64_t step_counter) {
    if (!run) return fail(AMH_ERR_INVALID, "run is NULL");
    amh_run& r = *run;
    const long long n = r.n, np = r.pitch;
    const int d = r.dim;
    if (ld < n) return fail(AMH_ERR_INVALID, "ld must be >= nchains_local");
    if (grad && !r.G) return fail(AMH_ERR_INVALID, "sampler keeps no gradient");
    if (S && !r.S) return fail(AMH_ERR_INVALID, "sampler keeps no Cholesky factor");
    AMH_CUDA_TRY(cudaSetDevice(r.ctx->device));
    cudaStream_t st = r.ctx->stream;
    if (x) AMH_CUDA_TRY(cudaMemcpy2DAsync(r.X, sizeof(double) * np, x, sizeof(double) * ld, sizeof(double) * n, d, cudaMemcpyHostToDevice, st));
    if (lp) AMH_CUDA_TRY(cudaMemcpyAsync(r.lp, lp, sizeof(double) * n, cudaMemcpyHostToDevice, st));
    if (grad) AMH_CUDA_TRY(cudaMemcpy2DAsync(r.G, sizeof(double) * np, grad, sizeof(double) * ld, sizeof(double) * n, d, cudaMemcpyHostToDevice, st));
    if (accepted) AMH_CUDA_TRY(cudaMemcpyAsync(r.acc, accepted, (size_t)n, cudaMemcpyHostToDevice, st));
    if (naccept) AMH_CUDA_TRY(cudaMemcpyAsync(r.nacc, naccept, sizeof(int64_t) * n, cudaMemcpyHostToDevice, st));
    int rc = AMH_OK;
    if (S) {
        const size_t nt = (size_t)d * (d + 1) / 2;
        double* tmp = nullptr;
        rc = dmalloc(r.ctx, (void**)&tmp, sizeof(double) * nt * np);
        if (rc) return rc;
        cudaError_t e = cudaMemcpy2DAsync(tmp, sizeof(double) * np, S, sizeof(double) * ld, sizeof(double) * n, nt, cudaMemcpyHostToDevice, st);
        if (e != cudaSuccess) rc = cuda_fail(e, "copy S");
        if (!rc) rc = r.ram_warp ? ramw_import_S(r, tmp) : ram_scatter_S(r, tmp);
        dfree(r.ctx, tmp);
        if (rc) return rc;
    }
    if (x && r.sampler->d.kind == AMH_SAMPLER_STATIC && !r.sampler->d.symmetric && !r.sampler->by_components()) {
        rc = launch_relq(r);
        if (rc) return rc;
    }
    if (step_counter >= 0) r.step = step_counter;
    AMH_CUDA_TRY(cudaStreamSynchronize(st));      /* caller buffers are only read during the call */
    return AMH_OK;
}

int32_t amh_run_get_ram_adapt(amh_run* run, double* logalpha, double* eta) {
    if (!run) return fail(AMH_ERR_INVALID, "run is NULL");
    amh_run& r = *run;
    if (r.sampler->d.kind != AMH_SAMPLER_RAM) return fail(AMH_ERR_INVALID, "not a RobustAdaptiveMetropolis run");
    AMH_CUDA_TRY(cudaSetDevice(r.ctx->device));
    AMH_CUDA_TRY(cudaStreamSynchronize(r.ctx->stream));
    if (logalpha) AMH_CUDA_TRY(cudaMemcpy(logalpha, r.logalpha, sizeof(double) * r.n, cudaMemcpyDeviceToHost));
    if (eta) AMH_CUDA_TRY(cudaMemcpy(eta, r.eta, sizeof(double) * r.n, cudaMemcpyDeviceToHost));
    return AMH_OK;
}

int32_t amh_run_set_ram_adapt(amh_run* run, const double* logalpha, const double* eta, const uint8_t* failed) {
    if (!run) return fail(AMH_ERR_INVALID, "run is NULL");
    amh_run& r = *run;
    if (r.sampler->d.kind != AMH_SAMPLER_RAM) return fail(AMH_ERR_INVALID, "not a RobustAdaptiveMetropolis run");
    AMH_CUDA_TRY(cudaSetDevice(r.ctx->device));
    cudaStream_t st = r.ctx->stream;
    if (logalpha) AMH_CUDA_TRY(cudaMemcpyAsync(r.logalpha, logalpha, sizeof(double) * r.n, cudaMemcpyHostToDevice, st));
    if (eta) AMH_CUDA_TRY(cudaMemcpyAsync(r.eta, eta, sizeof(double) * r.n, cudaMemcpyHostToDevice, st));
    if (failed) AMH_CUDA_TRY(cudaMemcpyAsync(r.failed, failed, (size_t)r.n, cudaMemcpyHostToDevice, st));
    AMH_CUDA_TRY(cudaStreamSynchronize(st));
    return AMH_OK;
}

int32_t amh_run_ram_failed(amh_run* run, int64_t* nfailed, int64_t* first_chain, uint8_t* failed) {
    if (!run) return fail(AMH_ERR_INVALID, "run is NULL");
    amh_run& r = *run;
    if (r.sampler->d.kind != AMH_SAMPLER_RAM) return fail(AMH_ERR_INVALID, "not a RobustAdaptiveMetropolis run");
    AMH_CUDA_TRY(cudaSetDevice(r.ctx->device));
    AMH_CUDA_TRY(cudaStreamSynchronize(r.ctx->stream));
    std::vector<unsigned char> h((size_t)r.n);
    AMH_CUDA_TRY(cudaMemcpy(h.data(), r.failed, (size_t)r.n, cudaMemcpyDeviceToHost));
    long long nf = 0, first = -1;
    for (long long c = 0; c < r.n; ++c)
        if (h[c]) { if (first < 0) first = c; ++nf; }
    if (nfailed) *nfailed = nf;
    if (first_chain) *first_chain = first < 0 ? -1 : r.off + first;
    if (failed) std::memcpy(failed, h.data(), (size_t)r.n);
    return AMH_OK;
}

int32_t amh_run_set_params(amh_run* run, const double* x) {
    if (!run || !x) return fail(AMH_ERR_INVALID, "NULL argument");
    amh_run& r = *run;
    AMH_CUDA_TRY(cudaSetDevice(r.ctx->device));
    AMH_CUDA_TRY(cudaMemcpy2DAsync(r.X, sizeof(double) * r.pitch, x, sizeof(double) * r.n, sizeof(double) * r.n, r.dim,
                                   cudaMemcpyHostToDevice, r.ctx->stream));
    int rc = AMH_OK;
    /* setparams!! recomputes lp (src/AdvancedMH.jl:151-157) and the gradient (MALA.jl:27-35);
     * RAM's setparams!! keeps logprob and S (RobustAdaptiveMetropolis.jl:117-121) */
    if (r.sampler->d.kind != AMH_SAMPLER_RAM) {
        double* Ssave = r.S;
        r.S = nullptr;
        r.keep_acc = true;               /* setparams!! builds Transition(model, params, t.accepted): the flag is kept */
        rc = launch_init(r, 0);
        r.keep_acc = false;
        r.S = Ssave;
    }
    AMH_CUDA_TRY(cudaStreamSynchronize(r.ctx->stream));
    return rc;
}

int32_t amh_host_alloc(size_t bytes, void** out) {
    if (!out) return fail(AMH_ERR_INVALID, "out is NULL");
    *out = nullptr;
    if (bytes == 0) return AMH_OK;
    AMH_CUDA_TRY(cudaHostAlloc(out, bytes, cudaHostAllocPortable));     /* pinned for every device of a multi-GPU job */
    return AMH_OK;
}
int32_t amh_host_free(void* p) {
    if (p) AMH_CUDA_TRY(cudaFreeHost(p));
    return AMH_OK;
}

int32_t amh_run_dim(amh_run* run) { return run ? run->dim : -1; }
int32_t amh_run_contract(amh_run* run) { return run ? run->cv : -1; }
int64_t amh_run_nchains(amh_run* run) { return run ? run->n : -1; }
int64_t amh_run_launch_count(amh_run* run) { return run ? run->launches : -1; }

int32_t amh_run_kernel_time_ms(amh_run* run, int32_t reset, double* ms, int64_t* launches) {
    if (!run) return fail(AMH_ERR_INVALID, "run is NULL");
    AMH_CUDA_TRY(cudaSetDevice(run->ctx->device));
    run->timing = true;           /* from now on step launches are bracketed by CUDA events */
    const int rc = resolve_events(*run);
    if (rc) return rc;
    if (ms) *ms = run->kernel_ms;
    if (launches) *launches = run->timed_launches;
    if (reset) { run->kernel_ms = 0; run->timed_launches = 0; }
    return AMH_OK;
}

}  /* extern "C" */
