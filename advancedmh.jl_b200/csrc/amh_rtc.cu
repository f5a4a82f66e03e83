/* amh_rtc.cu -- user-supplied targets: run-time compilation of the generic kernels with the
 * caller's log-density source (amh_target_create_source in include/amh.h).
 *
 * What stands behind it in the reference: DensityModel(f) wraps ANY Julia closure (src/AdvancedMH.jl:52-54,74) and
 * LogDensityModel wraps any LogDensityProblems object (src/AdvancedMH.jl:76, MALA.jl:100-105).  A closure cannot cross
 * a C ABI into a kernel, so the function arrives as C++ source text and is compiled here, on the host that owns the
 * GPU, by NVRTC:
 *   - the translation unit is this library's own device code (numerical contract, proposal algebra, kernels K1, K1C,
 *     K2/K2F, K3, K4, K6), embedded at build time by tools/gen_rtc_source.py, with the user's source appended;
 *   - options: sm_100a, C++17, --fmad=false (the contract: every fused multiply-add is an explicit fma());
 *   - one small module per kernel, compiled the first time a sampler needs it (the first-step kernel is compiled at
 *     target creation so that errors in the user's source surface there, with the compiler log);
 *   - libnvrtc is dlopen'ed and the driver entry points come from cudaGetDriverEntryPoint, so libamh_b200.so has no
 *     link-time dependency on either.
 * There is no interpreter and no host evaluation: a target that fails to compile is an error. */
#include <cuda.h>
#include <dlfcn.h>
#include <nvrtc.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include "amh_host.h"
#include "amh_rtc_source.inc"

namespace amhh {

namespace {

struct Nvrtc {
    void* h = nullptr;
    decltype(&nvrtcCreateProgram) CreateProgram = nullptr;
    decltype(&nvrtcDestroyProgram) DestroyProgram = nullptr;
    decltype(&nvrtcCompileProgram) CompileProgram = nullptr;
    decltype(&nvrtcGetCUBINSize) GetCUBINSize = nullptr;
    decltype(&nvrtcGetCUBIN) GetCUBIN = nullptr;
    decltype(&nvrtcGetProgramLogSize) GetProgramLogSize = nullptr;
    decltype(&nvrtcGetProgramLog) GetProgramLog = nullptr;
    decltype(&nvrtcAddNameExpression) AddNameExpression = nullptr;
    decltype(&nvrtcGetLoweredName) GetLoweredName = nullptr;
    decltype(&nvrtcGetErrorString) GetErrorString = nullptr;
    decltype(&nvrtcVersion) Version = nullptr;
    int major = 0, minor = 0;
    std::string why;
    bool has_v4_f64() const { return major > 12 || (major == 12 && minor >= 9); }   /* 256-bit ld/st.global in PTX */
};

struct Driver {
    decltype(&cuModuleLoadData) ModuleLoadData = nullptr;
    decltype(&cuModuleUnload) ModuleUnload = nullptr;
    decltype(&cuModuleGetFunction) ModuleGetFunction = nullptr;
    decltype(&cuFuncSetAttribute) FuncSetAttribute = nullptr;
    decltype(&cuLaunchKernel) LaunchKernel = nullptr;
    decltype(&cuGetErrorString) GetErrorString = nullptr;
    bool ok = false;
};

Nvrtc& nvrtc() {
    static Nvrtc n;
    static std::once_flag once;
    std::call_once(once, [] {
        /* the toolkit's own library first, by absolute path: another component of the process (PyTorch) may already
         * have loaded an OLDER libnvrtc.so.12 under the same soname, and a bare name would return that one */
        const char* names[] = {"/usr/local/cuda/lib64/libnvrtc.so.12", "/usr/local/cuda/targets/x86_64-linux/lib/libnvrtc.so.12",
                               "libnvrtc.so.12", "libnvrtc.so"};
        std::string tried;
        void* best = nullptr;
        int bmaj = 0, bmin = 0;
        auto consider = [&](const char* name) {
            void* h = dlopen(name, RTLD_NOW | RTLD_LOCAL);
            if (!h) { tried += std::string(" ") + name; return; }
            auto ver = (decltype(&nvrtcVersion))dlsym(h, "nvrtcVersion");
            int mj = 0, mn = 0;
            if (ver) ver(&mj, &mn);
            if (!best || mj > bmaj || (mj == bmaj && mn > bmin)) {
                if (best) dlclose(best);
                best = h; bmaj = mj; bmin = mn;
            } else {
                dlclose(h);
            }
        };
        if (const char* ev = std::getenv("AMH_NVRTC_LIB")) consider(ev);
        for (size_t i = 0; i < sizeof(names) / sizeof(names[0]) && !(bmaj > 12 || (bmaj == 12 && bmin >= 9)); ++i) consider(names[i]);
        n.h = best; n.major = bmaj; n.minor = bmin;
        if (!n.h) { n.why = "libnvrtc not found (tried" + tried + "; set AMH_NVRTC_LIB)"; return; }
#define AMH_NVRTC_SYM(name)                                                             \
        n.name = (decltype(n.name))dlsym(n.h, "nvrtc" #name);                          \
        if (!n.name) { n.why = "libnvrtc lacks nvrtc" #name; dlclose(n.h); n.h = nullptr; return; }
        AMH_NVRTC_SYM(CreateProgram) AMH_NVRTC_SYM(DestroyProgram) AMH_NVRTC_SYM(CompileProgram)
        AMH_NVRTC_SYM(GetCUBINSize) AMH_NVRTC_SYM(GetCUBIN) AMH_NVRTC_SYM(GetProgramLogSize)
        AMH_NVRTC_SYM(GetProgramLog) AMH_NVRTC_SYM(AddNameExpression) AMH_NVRTC_SYM(GetLoweredName)
        AMH_NVRTC_SYM(GetErrorString)
#undef AMH_NVRTC_SYM
    });
    return n;
}

Driver& driver() {
    static Driver d;
    static std::once_flag once;
    std::call_once(once, [] {
        bool ok = true;
#define AMH_DRV_SYM(field, sym)                                                                                   \
        {                                                                                                           \
            void* p = nullptr;                                                                                      \
            cudaDriverEntryPointQueryResult q;                                                                      \
            if (cudaGetDriverEntryPoint(sym, &p, cudaEnableDefault, &q) != cudaSuccess || !p) ok = false;           \
            d.field = (decltype(d.field))p;                                                                         \
        }
        AMH_DRV_SYM(ModuleLoadData, "cuModuleLoadData") AMH_DRV_SYM(ModuleUnload, "cuModuleUnload")
        AMH_DRV_SYM(ModuleGetFunction, "cuModuleGetFunction") AMH_DRV_SYM(FuncSetAttribute, "cuFuncSetAttribute")
        AMH_DRV_SYM(LaunchKernel, "cuLaunchKernel") AMH_DRV_SYM(GetErrorString, "cuGetErrorString")
#undef AMH_DRV_SYM
        d.ok = ok;
    });
    return d;
}

int drv_fail(CUresult e, const char* what) {
    const char* s = nullptr;
    if (driver().GetErrorString) driver().GetErrorString(e, &s);
    return fail(AMH_ERR_CUDA, std::string(what) + ": " + (s ? s : "CUDA driver error"));
}

/* the instantiation each kernel id stands for; must match the template arguments the launchers assume */
const char* kernel_expr(int which) {
    switch (which) {
    case RK_INIT: return "&amhd::init_kernel<amhd::TUser>";
    case RK_MH: return "&amhd::mh_step_kernel<0, amhd::TUser, 64, 4>";
    case RK_COMP: return "&amhh::mh_comp_kernel<amhd::TUser, 64>";
    case RK_MALA: return "&amhh::mala_step_kernel<0, amhd::TUser, 32>";
    case RK_RAM128: return "&amhh::ram_step_kernel<amhd::TUser, 128>";
    case RK_RAM64: return "&amhh::ram_step_kernel<amhd::TUser, 64>";
    case RK_RAM32: return "&amhh::ram_step_kernel<amhd::TUser, 32>";
    case RK_STRETCH: return "&amhh::stretch_sweep_kernel<0, amhd::TUser, 1024>";
    case RK_FLOW512: return "&amhh::stretch_sweep_flow_kernel<0, amhd::TUser, 512>";
    case RK_FLOW768: return "&amhh::stretch_sweep_flow_kernel<0, amhd::TUser, 768>";
    case RK_FLOW1024: return "&amhh::stretch_sweep_flow_kernel<0, amhd::TUser, 1024>";
    }
    return nullptr;
}

}  // namespace

struct RtcKernelSlot {
    CUmodule mod = nullptr;
    CUfunction fn = nullptr;
    int max_smem = 0;
};

struct RtcState {
    std::string tu;                     /* library device code + user source */
    RtcKernelSlot slot[RK_COUNT];
};

static int compile_kernel(amh_target& t, int which) {
    RtcState& st = *t.rtc;
    Nvrtc& n = nvrtc();
    if (!n.h) return fail(AMH_ERR_UNSUPPORTED, "user-supplied targets need NVRTC: " + n.why);
    Driver& d = driver();
    if (!d.ok) return fail(AMH_ERR_CUDA, "CUDA driver entry points unavailable (cudaGetDriverEntryPoint)");
    AMH_CUDA_TRY(cudaSetDevice(t.ctx->device));
    AMH_CUDA_TRY(cudaFree(nullptr));     /* primary context current on this thread */
    nvrtcProgram prog = nullptr;
    nvrtcResult rc = n.CreateProgram(&prog, st.tu.c_str(), "amh_user_target.cu", 0, nullptr, nullptr);
    if (rc != NVRTC_SUCCESS) return fail(AMH_ERR_CUDA, std::string("nvrtcCreateProgram: ") + n.GetErrorString(rc));
    const char* expr = kernel_expr(which);
    n.AddNameExpression(prog, expr);
    const char* opts[] = {"--gpu-architecture=sm_100a", "--std=c++17", "--fmad=false", "--generate-line-info",
                          "--prec-div=true", "--prec-sqrt=true", "--ftz=false", "--diag-suppress=177"};
    rc = n.CompileProgram(prog, (int)(sizeof(opts) / sizeof(opts[0])), opts);
    size_t lsz = 0;
    n.GetProgramLogSize(prog, &lsz);
    std::string log(lsz, '\0');
    if (lsz > 1) n.GetProgramLog(prog, &log[0]);
    while (!log.empty() && (log.back() == '\0' || log.back() == '\n')) log.pop_back();
    t.build_log = log;
    if (rc != NVRTC_SUCCESS) {
        n.DestroyProgram(&prog);
        return fail(AMH_ERR_INVALID, std::string("the target source does not compile (") + n.GetErrorString(rc) + "):\n" + log);
    }
    const char* lowered = nullptr;
    rc = n.GetLoweredName(prog, expr, &lowered);
    if (rc != NVRTC_SUCCESS || !lowered) {
        n.DestroyProgram(&prog);
        return fail(AMH_ERR_CUDA, std::string("nvrtcGetLoweredName failed for ") + expr);
    }
    size_t csz = 0;
    n.GetCUBINSize(prog, &csz);
    std::vector<char> cubin(csz);
    rc = n.GetCUBIN(prog, cubin.data());
    const std::string lname(lowered);
    n.DestroyProgram(&prog);
    if (rc != NVRTC_SUCCESS || csz == 0) return fail(AMH_ERR_CUDA, "nvrtcGetCUBIN failed");
    RtcKernelSlot& k = st.slot[which];
    CUresult e = d.ModuleLoadData(&k.mod, cubin.data());
    if (e != CUDA_SUCCESS) return drv_fail(e, "cuModuleLoadData");
    e = d.ModuleGetFunction(&k.fn, k.mod, lname.c_str());
    if (e != CUDA_SUCCESS) return drv_fail(e, "cuModuleGetFunction");
    return AMH_OK;
}

int rtc_build(amh_target& t, const char* source, bool has_grad) {
    t.rtc = new RtcState();
    size_t total = 0;
    for (int i = 0; i < kRtcSourceChunkCount; ++i) total += std::strlen(kRtcSourceChunks[i]);
    t.rtc->tu.reserve(total + std::strlen(source) + 64);
    /* without a gradient the kernels never reference amh_user_logdensity_and_gradient (amh_device.cuh, TUser) */
    if (!has_grad) t.rtc->tu += "#define AMH_RTC_NO_GRADIENT 1\n";
    /* ld/st.global.v4.f64 (256-bit) needs the ptxas of CUDA 12.9; an older NVRTC gets two 128-bit accesses instead */
    if (nvrtc().h && !nvrtc().has_v4_f64()) t.rtc->tu += "#define AMH_NO_V4_F64 1\n";
    for (int i = 0; i < kRtcSourceChunkCount; ++i) t.rtc->tu += kRtcSourceChunks[i];
    t.rtc->tu += source;
    t.rtc->tu += "\n";
    /* the first-step kernel uses both entry points: compiling it now reports errors in the user's source here */
    return compile_kernel(t, RK_INIT);
}

void rtc_destroy(amh_target& t) {
    if (!t.rtc) return;
    Driver& d = driver();
    for (int i = 0; i < RK_COUNT; ++i)
        if (t.rtc->slot[i].mod && d.ok) d.ModuleUnload(t.rtc->slot[i].mod);
    delete t.rtc;
    t.rtc = nullptr;
}

int rtc_launch(amh_run& r, int which, unsigned grid, unsigned block, size_t smem, void** params) {
    amh_target& t = *r.target;
    if (!t.rtc) return fail(AMH_ERR_STATE, "target has no run-time module");
    RtcKernelSlot& k = t.rtc->slot[which];
    if (!k.fn) {
        const int rc = compile_kernel(t, which);
        if (rc) return rc;
    }
    Driver& d = driver();
    if ((int)smem > 48 * 1024 && (int)smem > k.max_smem) {
        const CUresult e = d.FuncSetAttribute(k.fn, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, (int)smem);
        if (e != CUDA_SUCCESS) return drv_fail(e, "cuFuncSetAttribute(max dynamic shared memory)");
        k.max_smem = (int)smem;
    }
    const CUresult e = d.LaunchKernel(k.fn, grid, 1, 1, block, 1, 1, (unsigned)smem, (CUstream)r.ctx->stream, params, nullptr);
    if (e != CUDA_SUCCESS) return drv_fail(e, "cuLaunchKernel");
    return AMH_OK;
}

}  // namespace amhh
