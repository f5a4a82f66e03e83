// issue_probe.cu -- do integer issue slots overlap with a busy FP64 datapath on B200?
// One CTA of 8 warps per SM (2 warps per SM sub-partition).  Warps 0..3 run stream A, warps 4..7 stream B
// (warp w and w+4 share a sub-partition).  Each role is timed alone (the other half exits) and together.
//   streams: DMMA (4 independent accumulators), DFMA (8 chains), INT (8 chains of IMAD.WIDE + LOP3, Philox-like),
//            LDS (8 independent shared-memory loads per iteration)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o issue_probe issue_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

enum { S_NONE = 0, S_DMMA = 1, S_DFMA = 2, S_INT = 3, S_LDS = 4, S_HILO = 5, S_LO = 6, S_HI = 7, S_LOP = 8 };

template <int KIND>
__device__ __forceinline__ double stream(int iters, double a, double b, const double* sm) {
    double acc = 0;
    if (KIND == S_DMMA) {
        double c[4][2];
#pragma unroll
        for (int i = 0; i < 4; ++i) { c[i][0] = i; c[i][1] = -i; }
        for (int it = 0; it < iters; ++it)
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int i = 0; i < 4; ++i) dmma(c[i][0], c[i][1], a, b);
#pragma unroll
        for (int i = 0; i < 4; ++i) acc += c[i][0] + c[i][1];
    } else if (KIND == S_DFMA) {
        double f[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) f[i] = threadIdx.x * 1e-3 + i;
        for (int it = 0; it < iters; ++it)
#pragma unroll
            for (int r = 0; r < 8; ++r)
#pragma unroll
                for (int i = 0; i < 8; ++i) f[i] = fma(f[i], a, b);
#pragma unroll
        for (int i = 0; i < 8; ++i) acc += f[i];
    } else if (KIND == S_INT) {
        unsigned x[8], y[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) { x[i] = threadIdx.x + i; y[i] = 17u * i; }
        for (int it = 0; it < iters; ++it)
#pragma unroll
            for (int r = 0; r < 8; ++r)
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const unsigned long long p = (unsigned long long)0xD2511F53u * x[i];
                    const unsigned n = (unsigned)(p >> 32) ^ y[i] ^ 0x9E3779B9u;
                    y[i] = (unsigned)p; x[i] = n;
                }
#pragma unroll
        for (int i = 0; i < 8; ++i) acc += (double)(x[i] ^ y[i]);
    } else if (KIND == S_HILO || KIND == S_LO || KIND == S_HI || KIND == S_LOP) {
        unsigned x[8], y[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) { x[i] = threadIdx.x + i; y[i] = 17u * i; }
        for (int it = 0; it < iters; ++it)
#pragma unroll
            for (int r = 0; r < 8; ++r)
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    unsigned hi = x[i], lo = y[i];
                    if (KIND == S_HILO || KIND == S_HI) asm volatile("mul.hi.u32 %0, %1, %2;" : "=r"(hi) : "r"(x[i]), "r"(0xD2511F53u));
                    if (KIND == S_HILO || KIND == S_LO) asm volatile("mul.lo.u32 %0, %1, %2;" : "=r"(lo) : "r"(x[i]), "r"(0xD2511F53u));
                    if (KIND == S_LOP) { hi = (x[i] >> 3) ^ y[i]; lo = x[i] + 0x9E3779B9u; }
                    const unsigned n = hi ^ y[i] ^ 0x9E3779B9u;
                    y[i] = lo; x[i] = n;
                }
#pragma unroll
        for (int i = 0; i < 8; ++i) acc += (double)(x[i] ^ y[i]);
    } else if (KIND == S_LDS) {
        double f[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) f[i] = 0;
        int o = threadIdx.x & 31;
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int r = 0; r < 8; ++r)
#pragma unroll
                for (int i = 0; i < 8; ++i) f[i] += sm[(o + 32 * i + 256 * r) & 2047];
            o = (o + 1) & 31;
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) acc += f[i];
    }
    return acc;
}

template <int KA, int KB>
__global__ void __launch_bounds__(256) probe(double* out, long long* cyc, int iters, double a, double b, int mode) {
    __shared__ double sm[2048];
    for (int i = threadIdx.x; i < 2048; i += 256) sm[i] = i * 1e-9;
    __syncthreads();
    const int warp = threadIdx.x >> 5;
    const bool roleA = warp < 4;
    if ((mode == 1 && !roleA) || (mode == 2 && roleA)) return;
    const long long t0 = clock64();
    double r = roleA ? stream<KA>(iters, a, b, sm) : stream<KB>(iters, a, b, sm);
    const long long t1 = clock64();
    out[blockIdx.x * 256 + threadIdx.x] = r;
    if ((threadIdx.x & 31) == 0) cyc[blockIdx.x * 8 + warp] = t1 - t0;
}

template <int KA, int KB>
void run(const char* name, int iters) {
    int sms; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    double* out; long long* cyc;
    CK(cudaMalloc(&out, sizeof(double) * sms * 256));
    CK(cudaMallocManaged(&cyc, sizeof(long long) * sms * 8));
    double res[3][2];
    for (int mode = 0; mode < 3; ++mode) {
        for (int i = 0; i < sms * 8; ++i) cyc[i] = 0;
        probe<KA, KB><<<sms, 256>>>(out, cyc, iters, 0.999, 1e-3, mode);
        CK(cudaDeviceSynchronize());
        double sa = 0, sb = 0;
        for (int c = 0; c < sms; ++c)
            for (int w = 0; w < 8; ++w) (w < 4 ? sa : sb) += (double)cyc[c * 8 + w];
        res[mode][0] = sa / (sms * 4.0) / iters;
        res[mode][1] = sb / (sms * 4.0) / iters;
    }
    printf("%-14s cycles/iter  A alone %8.1f  B alone %8.1f | together A %8.1f (x%.2f)  B %8.1f (x%.2f)\n", name,
           res[1][0], res[2][1], res[0][0], res[0][0] / res[1][0], res[0][1], res[0][1] / res[2][1]);
    CK(cudaFree(out)); CK(cudaFree(cyc));
}

int main() {
    // per iteration: DMMA 16 instr, DFMA 64 instr, INT 64 x (IMAD.WIDE + LOP3 + ...), LDS 64 loads
    run<S_DMMA, S_INT>("DMMA | INT", 4000);
    run<S_DFMA, S_INT>("DFMA | INT", 4000);
    run<S_DMMA, S_DFMA>("DMMA | DFMA", 4000);
    run<S_DMMA, S_LDS>("DMMA | LDS", 4000);
    run<S_DFMA, S_LDS>("DFMA | LDS", 4000);
    run<S_INT, S_INT>("INT  | INT", 4000);
    run<S_DMMA, S_HILO>("DMMA | HI+LO", 4000);
    run<S_DMMA, S_HI>("DMMA | HI", 4000);
    run<S_DMMA, S_LO>("DMMA | LO", 4000);
    run<S_DMMA, S_LOP>("DMMA | LOP", 4000);
    run<S_HILO, S_HILO>("HI+LO | HI+LO", 4000);
    run<S_HI, S_HI>("HI | HI", 4000);
    run<S_LO, S_LO>("LO | LO", 4000);
    run<S_LOP, S_LOP>("LOP | LOP", 4000);
    run<S_DMMA, S_DMMA>("DMMA | DMMA", 4000);
    run<S_DFMA, S_DFMA>("DFMA | DFMA", 4000);
    return 0;
}
