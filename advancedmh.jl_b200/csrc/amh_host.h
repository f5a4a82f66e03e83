/* amh_host.h -- host-side objects behind the opaque handles of include/amh.h. */
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <set>
#include <string>
#include <vector>
#include "../../include/amh.h"
#include "amh_device.cuh"

struct amh_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t copy_stream = nullptr;
    cudaMemPool_t pool = nullptr;         /* this context's stream-ordered memory pool */
    int sm_count = 0;
    std::set<const void*> configured;     /* kernels whose function attributes were set on this device */
    /* How host threads wait for this context's streams.  false (default): cudaStreamSynchronize -- the driver spins,
     * lowest latency.  true: sleep on an event created with cudaEventBlockingSync (AMH_JOB_BLOCKING=1 sets it for the
     * contexts of a multi-device job: for hosts with fewer free cores than devices; ~0.3 ms of wake-up latency per wait
     * on the 2 x B200 box, profiles/r2_job_fanout_2gpu.txt). */
    bool blocking_wait = false;
    cudaEvent_t wait_ev = nullptr;
};

namespace amhh { struct RtcState; }

struct amh_target {
    amh_ctx* ctx = nullptr;
    int kind = 0, dim = 0;
    std::vector<double> blob;      /* host copy */
    double* dblob = nullptr;       /* device copy of the whole blob */
    long long ndata = 0;
    double inv2tau2 = 0, invtau2 = 0;
    /* AMH_TARGET_USER: run-time compiled kernels (amh_rtc.cu) */
    amhh::RtcState* rtc = nullptr;
    bool user_grad = false;
    std::string build_log;
    bool has_grad() const {
        if (kind == AMH_TARGET_USER) return user_grad;
        return kind == AMH_TARGET_MVNORMAL || kind == AMH_TARGET_GAUSS_PREC || kind == AMH_TARGET_IID_NORMAL ||
               kind == AMH_TARGET_LOGISTIC || kind == AMH_TARGET_ROSENBROCK;
    }
};

struct amh_sampler {
    amh_ctx* ctx = nullptr;
    amh_sampler_desc d{};
    std::vector<double> mean, scale, S0;   /* host copies (S0 packed lower) */
    bool has_mean = false;
    double mala_sigma = 0;
    double* dmean = nullptr;
    double* dscale = nullptr;
    double* dS0 = nullptr;
    std::vector<amh_component> comps;      /* AMH_COV_COMPONENTS / AMH_SAMPLER_MIXED: one univariate law per coordinate */
    amh_component* dcomps = nullptr;
    bool by_components() const { return !comps.empty(); }
};

struct amh_run {
    amh_ctx* ctx = nullptr;
    amh_target* target = nullptr;
    amh_sampler* sampler = nullptr;
    long long n = 0, off = 0, pitch = 0;
    int dim = 0;
    int x_rows = 0;                /* rows X was allocated with: dim rounded up to a multiple of 8, the extra rows stay 0 */
    long long nseeds = 0;
    /* device state */
    double* X = nullptr;
    double* X2 = nullptr;          /* second walker buffer (stretch) / second S buffer flagging (RAM) */
    double* lp = nullptr;
    double* lp2 = nullptr;
    double* lq = nullptr;
    double* G = nullptr;
    double* S = nullptr;
    double* S2 = nullptr;
    double* logalpha = nullptr;
    double* eta = nullptr;
    unsigned char* acc = nullptr;
    unsigned char* failed = nullptr;
    unsigned char* sflag = nullptr;
    unsigned long long* nacc = nullptr;
    unsigned long long* seeds = nullptr;
    double* sum = nullptr;
    double* sumsq = nullptr;
    void* scratch = nullptr;       /* sampler / kernel specific device buffer */
    void* scratch2 = nullptr;      /* K3T: bf16 operand slices, candidate / gradient scratch */
    void* tensor_state = nullptr;  /* K3T: host-side state (tensor maps), amh_launch_mala_tensor.cu */
    size_t scratch_bytes = 0;
    long long step = 0;
    long long nsaved = 0;
    long long launches = 0;
    /* stretch move (K2F): the sweep plans are state independent, so the plan of the NEXT launch is computed on a second
     * stream while the current sweeps run (the sweep kernel occupies one SM per ensemble, the rest of the GPU is idle) */
    cudaStream_t aux_stream = nullptr;
    cudaEvent_t ev_plan[2] = {nullptr, nullptr};     /* plan in buffer b complete (recorded on the stream that made it) */
    cudaEvent_t ev_sweep[2] = {nullptr, nullptr};    /* last sweep kernel that read buffer b complete                   */
    long long plan_step0[2] = {-1, -1};              /* what buffer b holds: first step and number of sweeps            */
    int plan_nsteps[2] = {0, 0};
    int plan_layout_nsteps = -1;                     /* launch length the two plan buffers are laid out for            */
    int cv = AMH_CONTRACT_VERSION;  /* contract version of the step noise (amh_sampler_desc.contract) */
    bool keep_acc = false;         /* launch_init leaves Transition.accepted alone (amh_run_set_params) */
    bool ram_warp = false;         /* RAM: S stored [chain][column-packed] and stepped by K4W */
    int mh_path = 0;               /* 0 = choose (tensor-core K1T when eligible), 1 = force the per-thread DFMA kernel K1 */
    /* kernel timing */
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> pending;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> pool;
    double kernel_ms = 0;
    long long timed_launches = 0;
    long long pending_launches = 0;
    bool timing = false;           /* record CUDA events around step launches (enabled by amh_run_kernel_time_ms) */
};

namespace amhh {
inline int tri_h(int i, int j) { return i * (i + 1) / 2 + j; }
int fail(int code, const std::string& msg);
int cuda_fail(cudaError_t e, const char* what);
#define AMH_CUDA_TRY(expr)                                   \
    do {                                                     \
        cudaError_t _e = (expr);                             \
        if (_e != cudaSuccess) return amhh::cuda_fail(_e, #expr); \
    } while (0)

/* host wait for everything enqueued on `st` so far, spinning or sleeping as the context says */
cudaError_t sync_stream(amh_ctx* ctx, cudaStream_t st);
void ctx_set_blocking_wait(amh_ctx* ctx, bool on);
amhd::ChainState chain_state(amh_run& r);
int target_create_impl(amh_ctx* ctx, int32_t kind, int32_t dim, const double* blob, int64_t nblob, amh_target** out, bool upload_data);
int target_create_empty(amh_ctx* ctx, int32_t kind, int32_t dim, const double* blob, int64_t nblob, amh_target** out);
/* device memory from the context's stream-ordered pool (cudaMallocAsync): run handles are created and
 * destroyed per `sample` call, and the pool makes that cheap */
int dmalloc(amh_ctx* ctx, void** p, size_t bytes);
void dfree(amh_ctx* ctx, void* p);

/* per-sampler launchers; each enqueues kernels on r.ctx->stream and bumps r.launches */
int launch_mh(amh_run& r, int nsteps, const amhd::SaveArgs& sv);
int launch_mh_hast(amh_run& r, int nsteps, const amhd::SaveArgs& sv, bool& taken);        /* amh_launch_mh_hast.cu */
int launch_mh_more_dims(amh_run& r, int nsteps, const amhd::SaveArgs& sv, bool& taken);   /* amh_launch_mh_dims.cu */
/* K1C: arrays of univariate proposal laws / arrays of proposals (amh_launch_mh_comp.cu) */
int launch_mh_comp(amh_run& r, int nsteps, const amhd::SaveArgs& sv);
bool mh_tc_eligible(const amh_run& r);        /* K1T: both mat-vecs on the FP64 tensor cores */
int launch_mh_tc(amh_run& r, int nsteps, const amhd::SaveArgs& sv);
/* the same kernels for the dimensions in between, padded to a multiple of 8, d <= 64 (amh_launch_mh_tcp.cu) */
bool mh_tc_padded_eligible(const amh_run& r);
bool mh_tc_small_eligible(const amh_run& r);     /* few chains: the 4-warp CTA shape, served by launch_mh_tc_padded too */
int launch_mh_tc_padded(amh_run& r, int nsteps, const amhd::SaveArgs& sv);
int launch_mala(amh_run& r, int nsteps, const amhd::SaveArgs& sv);
int launch_mala_more_dims(amh_run& r, int nsteps, const amhd::SaveArgs& sv, bool& taken);   /* amh_launch_mala_dims.cu */
/* K3T: the same on the tcgen05 tensor cores as split-bf16 GEMMs, opt-in (amh_launch_mala_tensor.cu) */
bool mala_tensor_eligible(const amh_run& r);
int launch_mala_tensor(amh_run& r, int nsteps, const amhd::SaveArgs& sv);
void mala_tensor_release(amh_run& r);
/* K3L: MALA on the many-row logistic target as two chained DMMA GEMMs (amh_launch_mala_logistic.cu) */
bool mala_logistic_eligible(const amh_run& r);
int launch_mala_logistic(amh_run& r, int nsteps, const amhd::SaveArgs& sv);
/* the same kernel without the gradient: RWMH with an isotropic / diagonal proposal on the many-row logistic target */
bool mh_logistic_eligible(const amh_run& r);
int launch_mh_logistic(amh_run& r, int nsteps, const amhd::SaveArgs& sv);
int launch_ram(amh_run& r, int nsteps, bool warmup, const amhd::SaveArgs& sv);
int launch_stretch(amh_run& r, int nsteps, const amhd::SaveArgs& sv);
/* exact-dimension stretch kernels of the second translation unit (amh_launch_stretch_dims.cu); taken = false: not one of its dimensions */
int launch_stretch_more_dims(amh_run& r, int nsteps, const amhd::SaveArgs& sv, bool& taken);
int launch_init(amh_run& r, int mode);
int ram_gather_S(amh_run& r, double* dst);
int ram_scatter_S(amh_run& r, const double* src);   /* [tri][pitch] device buffer -> current factor of every chain */
int launch_relq(amh_run& r);                        /* static MH: recompute the cached logq(state) */
/* K4W: one warp per chain, factor resident in shared memory (amh_launch_ram_warp.cu) */
bool ram_warp_eligible(const amh_run& r);
int launch_ram_warp(amh_run& r, int nsteps, bool warmup, const amhd::SaveArgs& sv);
int ramw_init_S(amh_run& r);
int ramw_import_S(amh_run& r, const double* src);
int ramw_export_S(amh_run& r, double* dst);   /* current factor of every chain -> dst [tri][pitch] */
int default_steps_per_launch(const amh_run& r);
int stretch_default_steps_per_launch(const amh_run& r);   /* amh_launch_stretch.cu */
/* user-supplied targets: kernels compiled at run time by NVRTC (amh_rtc.cu) */
enum RtcKernel { RK_INIT = 0, RK_MH, RK_COMP, RK_MALA, RK_RAM128, RK_RAM64, RK_RAM32, RK_STRETCH, RK_FLOW512, RK_FLOW768,
                 RK_FLOW1024, RK_COUNT };
int rtc_build(amh_target& t, const char* source, bool has_grad);
void rtc_destroy(amh_target& t);
int rtc_launch(amh_run& r, int which, unsigned grid, unsigned block, size_t smem, void** params);
}  // namespace amhh
