// dmma_probe.cu -- B200 micro-benchmarks behind the K1 design decisions:
//   (1) DFMA issue rate, (2) mma.sync.m8n8k4.f64 (DMMA) rate, (3) both interleaved (do they share a pipe?),
//   (4) is DMMA bit-identical to a sequential IEEE fma chain over k (needed for the parity contract)?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o dmma_probe dmma_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b, double c0, double c1) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%4,%5};\n"
                 : "=d"(d0), "=d"(d1) : "d"(a), "d"(b), "d"(c0), "d"(c1));
}

template <int NDFMA, int NDMMA>
__global__ void rate_kernel(double* out, int iters, double a, double b) {
    double f[8], c[8][2];
#pragma unroll
    for (int i = 0; i < 8; ++i) { f[i] = threadIdx.x * 1e-3 + i; c[i][0] = i; c[i][1] = -i; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
#pragma unroll
            for (int i = 0; i < NDFMA; ++i) f[i] = fma(f[i], a, b);
#pragma unroll
            for (int i = 0; i < NDMMA; ++i) dmma(c[i][0], c[i][1], a, b, c[i][0], c[i][1]);
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += f[i] + c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NDFMA, int NDMMA>
void run_rate(const char* name, int warps_per_sm) {
    int sms; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    const int iters = 20000;
    const int threads = 32 * warps_per_sm;
    double* out; CK(cudaMalloc(&out, sizeof(double) * sms * threads));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    rate_kernel<NDFMA, NDMMA><<<sms, threads>>>(out, 100, 0.999, 1e-3);
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    rate_kernel<NDFMA, NDMMA><<<sms, threads>>>(out, iters, 0.999, 1e-3);
    CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize());
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    int clk_khz; CK(cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0));
    const double cycles = ms * 1e-3 * clk_khz * 1e3;
    const double dfma_w = (double)iters * 4 * NDFMA * warps_per_sm;   // warp-instructions per SM
    const double dmma_w = (double)iters * 4 * NDMMA * warps_per_sm;
    printf("%-28s warps/SM=%2d  ms=%8.3f  DFMA warp-inst/clk/SM=%6.3f (=%5.1f lanes)  DMMA warp-inst/clk/SM=%6.4f (=%6.1f FMA/clk/SM)  [clk %d MHz nominal]\n",
           name, warps_per_sm, ms, dfma_w / cycles, 32 * dfma_w / cycles, dmma_w / cycles, 256 * dmma_w / cycles, clk_khz / 1000);
    CK(cudaFree(out));
}

// ---- bit-exactness: D = A(8x4) * B(4x8) + C via DMMA vs sequential fma over k = 0..3
__global__ void exact_kernel(const double* A, const double* B, const double* C, double* D, int ntiles) {
    const int lane = threadIdx.x;
    for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const double a = A[t * 32 + (lane / 4) * 4 + lane % 4];          // A[row=lane/4][k=lane%4], row-major 8x4
        const double b = B[t * 32 + (lane % 4) * 8 + lane / 4];          // B[k=lane%4][n=lane/4],  row-major 4x8
        const double c0 = C[t * 64 + (lane / 4) * 8 + (lane % 4) * 2];   // C[row=lane/4][n=2*(lane%4)+{0,1}]
        const double c1 = C[t * 64 + (lane / 4) * 8 + (lane % 4) * 2 + 1];
        double d0, d1;
        dmma(d0, d1, a, b, c0, c1);
        D[t * 64 + (lane / 4) * 8 + (lane % 4) * 2] = d0;
        D[t * 64 + (lane / 4) * 8 + (lane % 4) * 2 + 1] = d1;
    }
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    printf("device %s sm_%d%d SMs=%d\n", p.name, p.major, p.minor, p.multiProcessorCount);
    for (int w : {4, 8, 16, 32}) {
        run_rate<8, 0>("DFMA only (8 chains)", w);
        run_rate<0, 8>("DMMA only (8 chains)", w);
        run_rate<8, 8>("DFMA+DMMA interleaved", w);
        run_rate<8, 2>("DFMA x8 + DMMA x2", w);
        run_rate<4, 1>("DFMA x4 + DMMA x1", w);
        run_rate<2, 0>("DFMA only (2 chains: latency)", w);
        run_rate<0, 1>("DMMA only (1 chain: latency)", w);
    }
    // exactness
    const int nt = 4096;
    std::vector<double> A(nt * 32), B(nt * 32), C(nt * 64), D(nt * 64);
    uint64_t s = 88172645463325252ull;
    auto rnd = [&]() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return (double)(int64_t)(s >> 11) / 9007199254740992.0 * 2 - 1; };
    for (auto& x : A) x = rnd() * exp(8 * rnd());
    for (auto& x : B) x = rnd() * exp(8 * rnd());
    for (auto& x : C) x = rnd() * exp(8 * rnd());
    for (int i = 0; i < 64 * 16; ++i) C[i] = 0.0;          // first tiles: C = 0 (matvec start)
    double *dA, *dB, *dC, *dD;
    CK(cudaMalloc(&dA, A.size() * 8)); CK(cudaMalloc(&dB, B.size() * 8)); CK(cudaMalloc(&dC, C.size() * 8)); CK(cudaMalloc(&dD, D.size() * 8));
    CK(cudaMemcpy(dA, A.data(), A.size() * 8, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, B.data(), B.size() * 8, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dC, C.data(), C.size() * 8, cudaMemcpyHostToDevice));
    exact_kernel<<<64, 32>>>(dA, dB, dC, dD, nt);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(D.data(), dD, D.size() * 8, cudaMemcpyDeviceToHost));
    long long bad_fwd = 0, bad_rev = 0, bad_pair = 0;
    for (int t = 0; t < nt; ++t)
        for (int r = 0; r < 8; ++r)
            for (int n = 0; n < 8; ++n) {
                double f = C[t * 64 + r * 8 + n], g = f;
                for (int k = 0; k < 4; ++k) f = fma(A[t * 32 + r * 4 + k], B[t * 32 + k * 8 + n], f);
                for (int k = 3; k >= 0; --k) g = fma(A[t * 32 + r * 4 + k], B[t * 32 + k * 8 + n], g);
                // pairwise: (a0b0 + a1b1) + (a2b2 + a3b3) + c  (one plausible non-sequential order)
                const double got = D[t * 64 + r * 8 + n];
                if (memcmp(&got, &f, 8)) ++bad_fwd;
                if (memcmp(&got, &g, 8)) ++bad_rev;
            }
    printf("DMMA vs sequential fma chain k=0..3 : %lld mismatches of %d\n", bad_fwd, nt * 64);
    printf("DMMA vs sequential fma chain k=3..0 : %lld mismatches of %d\n", bad_rev, nt * 64);
    (void)bad_pair;
    return 0;
}
