/* amh_job.cu -- the CUDA backend of the multi-GPU job layer (amh_job_impl.h): N B200s of one box driven from ONE
 * process, which is what a Julia host needs (`MCMCB200(ngpus = 8)`; north star: "chains shard across the 8 GPUs of one
 * box with only a one-time NCCL broadcast of the target's fixed data ... no per-step collective").
 *
 * Broadcast of the target blob (AMH_JOB_BCAST = auto | nccl | peer | h2d):
 *   nccl  host -> device 0, then ONE ncclBroadcast over NVLink / NVSwitch (communicators from ncclCommInitAll, created
 *         lazily at the first broadcast; libnccl.so.2 is dlopen'ed -- no link-time dependency, and inside a process
 *         that already holds an NCCL, e.g. PyTorch's, the loaded copy is reused)
 *   peer  host -> device 0, then cudaMemcpyPeerAsync 0 -> k over NVLink
 *   h2d   N host -> device copies, one per PCIe link, in parallel
 *   auto  (default) by MEASUREMENT on 8 x B200 (profiles/r2_job_check_8gpu.txt), BASELINE config 4's 10.3 MB blob:
 *         h2d 2.5 ms, nccl 7.1 ms (+ 1.8 s of one-time communicator set-up, 309 ms for the first call), peer 8.8 ms --
 *         every GPU has its own PCIe link, so N parallel host copies beat a staged fan-out until the blob is large
 *         enough for host-memory bandwidth to matter: h2d up to 256 MB, above that device 0 + NVLink peer copies.
 * Samples never cross NVLink: every device writes its column block of the caller's host array directly. */
#include <dlfcn.h>
#include <nccl.h>
#include <cstdio>
#include <cstdlib>
#include "amh_host.h"
#include "amh_job_impl.h"

namespace {

struct NcclApi {
    void* lib = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    bool ok() const { return lib && CommInitAll && CommDestroy && Broadcast && GroupStart && GroupEnd; }
};

NcclApi& nccl_api() {
    static NcclApi api = [] {
        NcclApi a;
        /* 1. a copy this process already holds (PyTorch's, the host application's): the loader would hand it out under the
         *    soname anyway, and two NCCLs in one process is what must not happen;
         * 2. AMH_NCCL_LIB (the Python host points it at the NCCL wheel next to PyTorch, so that a later `import torch`
         *    finds the version it was built against -- an older system libnccl loaded first made that import fail with an
         *    undefined ncclDevCommCreate);
         * 3. the system's. */
        a.lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL | RTLD_NOLOAD);
        const char* names[] = {std::getenv("AMH_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
        for (const char* nm : names) {
            if (a.lib) break;
            if (!nm || !*nm) continue;
            a.lib = dlopen(nm, RTLD_NOW | RTLD_LOCAL);
        }
        if (a.lib) {
            a.CommInitAll = (decltype(a.CommInitAll))dlsym(a.lib, "ncclCommInitAll");
            a.CommDestroy = (decltype(a.CommDestroy))dlsym(a.lib, "ncclCommDestroy");
            a.Broadcast = (decltype(a.Broadcast))dlsym(a.lib, "ncclBroadcast");
            a.GroupStart = (decltype(a.GroupStart))dlsym(a.lib, "ncclGroupStart");
            a.GroupEnd = (decltype(a.GroupEnd))dlsym(a.lib, "ncclGroupEnd");
            a.GetErrorString = (decltype(a.GetErrorString))dlsym(a.lib, "ncclGetErrorString");
        }
        return a;
    }();
    return api;
}

struct CudaBackend {
    struct Shared {
        std::vector<ncclComm_t> comms;       /* empty: no NCCL (one device, library missing, or not asked for) */
        std::string mode;                    /* nccl | peer | h2d */
        double comm_init_ms = 0;
    };
    static std::string last_error() { return amh_last_error(); }
    static int fail(int code, const std::string& m) { return amhh::fail(code, m); }
    static int ctx_create(int dev, amh_ctx** out) { return amh_ctx_create(dev, out); }
    static int ctx_destroy(amh_ctx* c) { return amh_ctx_destroy(c); }
    static int target_destroy(amh_target* t) { return amh_target_destroy(t); }
    static int target_create_source(amh_ctx* c, int32_t dim, const char* src, int32_t g, const double* data, int64_t nd, amh_target** out) {
        return amh_target_create_source(c, dim, src, g, data, nd, out);
    }
    static int sampler_create(amh_ctx* c, const amh_sampler_desc* d, amh_sampler** out) { return amh_sampler_create(c, d, out); }
    static int sampler_destroy(amh_sampler* s) { return amh_sampler_destroy(s); }
    static int run_create(amh_ctx* c, amh_target* t, amh_sampler* s, int64_t n, int64_t off, const uint64_t* seeds, const double* init,
                          int64_t ld, amh_run** out) { return amh_run_create(c, t, s, n, off, seeds, init, ld, out); }
    static int run_destroy(amh_run* r) { return amh_run_destroy(r); }
    static int run_steps(amh_run* r, int64_t n, int32_t w, int32_t spl) { return amh_run_steps(r, n, w, spl); }
    static int run_sync(amh_run* r) { return amh_run_sync(r); }
    static int run_sample_ld(amh_run* r, int64_t N, int64_t di, int64_t th, int64_t nw, double* out, int64_t old, uint8_t* acc, int64_t ald,
                             amh_summary* s) { return amh_run_sample_ld(r, N, di, th, nw, out, old, acc, ald, s); }
    static int run_get_state_ld(amh_run* r, int64_t ld, double* x, double* lp, double* g, double* S, uint8_t* a, int64_t* na, int64_t* st) {
        return amh_run_get_state_ld(r, ld, x, lp, g, S, a, na, st);
    }
    static int run_set_state_ld(amh_run* r, int64_t ld, const double* x, const double* lp, const double* g, const double* S, const uint8_t* a,
                                const int64_t* na, int64_t st) { return amh_run_set_state_ld(r, ld, x, lp, g, S, a, na, st); }
    static int run_get_ram_adapt(amh_run* r, double* la, double* eta) { return amh_run_get_ram_adapt(r, la, eta); }
    static int run_set_ram_adapt(amh_run* r, const double* la, const double* eta, const uint8_t* f) { return amh_run_set_ram_adapt(r, la, eta, f); }
    static int run_ram_failed(amh_run* r, int64_t* nf, int64_t* first, uint8_t* f) { return amh_run_ram_failed(r, nf, first, f); }
    static int64_t run_launch_count(amh_run* r) { return amh_run_launch_count(r); }
    static int run_contract(amh_run* r) { return amh_run_contract(r); }
    static int run_kernel_time_ms(amh_run* r, int32_t reset, double* ms, int64_t* l) { return amh_run_kernel_time_ms(r, reset, ms, l); }

    static int shared_init(amhjob::Job<CudaBackend>& j) {
        Shared& sh = j.shared;
        const char* ev = std::getenv("AMH_JOB_BCAST");
        std::string want = ev ? ev : "";
        if (j.ngpus == 1) { sh.mode = "h2d"; return AMH_OK; }
        /* AMH_JOB_BLOCKING=1: the worker threads SLEEP while they wait for their devices (amh_host.h, amh_ctx::blocking_wait)
         * instead of spinning inside the driver -- for hosts with fewer free cores than devices; costs wake-up latency
         * (~0.9 ms per sample() call on the 2 x B200 box, profiles/r2_job_fanout_2gpu.txt), so it is not the default */
        if (const char* bw = std::getenv("AMH_JOB_BLOCKING"))
            if (bw[0] == '1')
                for (int k = 0; k < j.ngpus; ++k) amhh::ctx_set_blocking_wait(j.ctx[k], true);
        /* peer access 0 <-> k: lets cudaMemcpyPeer go over NVLink without staging (harmless if already enabled) */
        for (int k = 1; k < j.ngpus; ++k) {
            int can = 0;
            if (cudaDeviceCanAccessPeer(&can, j.devices[k], j.devices[0]) == cudaSuccess && can) {
                cudaSetDevice(j.devices[k]);
                const cudaError_t e = cudaDeviceEnablePeerAccess(j.devices[0], 0);
                if (e != cudaSuccess) cudaGetLastError();                 /* cudaErrorPeerAccessAlreadyEnabled */
            }
        }
        if (want == "h2d" || want == "peer" || want == "nccl") sh.mode = want;
        else sh.mode = "auto";
        if (sh.mode == "nccl" && !nccl_api().ok())
            return fail(AMH_ERR_UNSUPPORTED, "AMH_JOB_BCAST=nccl but libnccl.so.2 could not be loaded");
        return AMH_OK;
    }
    /* communicators are only built when a broadcast really goes through NCCL (1.8 s for 8 GPUs) */
    static int nccl_init(amhjob::Job<CudaBackend>& j) {
        Shared& sh = j.shared;
        if (!sh.comms.empty()) return AMH_OK;
        NcclApi& api = nccl_api();
        const auto t0 = std::chrono::steady_clock::now();
        sh.comms.assign(j.ngpus, nullptr);
        const ncclResult_t r = api.CommInitAll(sh.comms.data(), j.ngpus, j.devices.data());
        sh.comm_init_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        if (r != ncclSuccess) {
            sh.comms.clear();
            return fail(AMH_ERR_CUDA, std::string("ncclCommInitAll failed: ") + (api.GetErrorString ? api.GetErrorString(r) : "?"));
        }
        return AMH_OK;
    }
    static void shared_destroy(amhjob::Job<CudaBackend>& j) {
        for (ncclComm_t c : j.shared.comms)
            if (c) nccl_api().CommDestroy(c);
        j.shared.comms.clear();
    }

    /* target blob: host -> device 0, then device 0 -> the others */
    static int target_broadcast(amhjob::Job<CudaBackend>& j, int32_t kind, int32_t dim, const double* blob, int64_t nblob) {
        Shared& sh = j.shared;
        const size_t bytes = sizeof(double) * (size_t)nblob;
        std::string mode = sh.mode;
        if (mode == "auto") mode = (j.ngpus == 1 || bytes <= (256ull << 20)) ? "h2d" : "peer";
        j.bcast_mode = mode;
        if (mode == "h2d" || nblob == 0)
            return j.each([&](int k) { return (int)amh_target_create(j.ctx[k], kind, dim, blob, nblob, &j.target[k]); });
        if (mode == "nccl")
            if (int rc0 = nccl_init(j)) return rc0;
        int rc = amh_target_create(j.ctx[0], kind, dim, blob, nblob, &j.target[0]);        /* validates, uploads, syncs */
        if (rc) return rc;
        for (int k = 1; k < j.ngpus && !rc; ++k) rc = amhh::target_create_empty(j.ctx[k], kind, dim, blob, nblob, &j.target[k]);
        if (rc) return rc;
        if (mode == "nccl") {
            NcclApi& api = nccl_api();
            ncclResult_t r = api.GroupStart();
            for (int k = 0; k < j.ngpus && r == ncclSuccess; ++k)
                r = api.Broadcast(j.target[0]->dblob, j.target[k]->dblob, (size_t)nblob, ncclDouble, 0, sh.comms[k], j.ctx[k]->stream);
            const ncclResult_t r2 = api.GroupEnd();
            if (r == ncclSuccess) r = r2;
            if (r != ncclSuccess)
                return fail(AMH_ERR_CUDA, std::string("ncclBroadcast failed: ") + (api.GetErrorString ? api.GetErrorString(r) : "?"));
        } else {
            for (int k = 1; k < j.ngpus; ++k) {
                AMH_CUDA_TRY(cudaSetDevice(j.devices[k]));
                AMH_CUDA_TRY(cudaMemcpyPeerAsync(j.target[k]->dblob, j.devices[k], j.target[0]->dblob, j.devices[0], bytes, j.ctx[k]->stream));
            }
        }
        for (int k = 0; k < j.ngpus; ++k) {
            AMH_CUDA_TRY(cudaSetDevice(j.devices[k]));
            AMH_CUDA_TRY(cudaStreamSynchronize(j.ctx[k]->stream));
        }
        return AMH_OK;
    }
};

}  // namespace

AMH_DEFINE_JOB_ABI(amh_, CudaBackend)

extern "C" double amh_job_comm_init_ms(amh_job* j) { return j ? ((amhjob::Job<CudaBackend>*)j)->shared.comm_init_ms : -1.0; }
