// fdiv_probe.cu -- are the branch-free division / square-root sequences used by K4W's Givens sweep
// (advancedmh.jl_b200/csrc/amh_fastmath.cuh) bit-identical to IEEE `/` and sqrt() on B200?
// Brute force on the device: random mantissas x random exponents inside the guarded range, plus the
// hard cases for Markstein's correction (quotients next to a rounding boundary: a = q*b for random q, b).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -I../../advancedmh.jl_b200/csrc -o fdiv_probe fdiv_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "amh_fastmath.cuh"

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ unsigned long long splitmix(unsigned long long& s) {
    s += 0x9E3779B97F4A7C15ull;
    unsigned long long z = s;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
__device__ __forceinline__ double rnd_double(unsigned long long& s, int emax) {
    const unsigned long long m = splitmix(s);
    const int e = (int)(splitmix(s) % (unsigned long long)(2 * emax + 1)) - emax;
    const unsigned long long bits = (m & 0x800FFFFFFFFFFFFFull) | ((unsigned long long)(1023 + e) << 52);
    return __longlong_as_double((long long)bits);
}

__global__ void probe(unsigned long long* bad, int iters) {
    unsigned long long s = 0x1234567ull + 0x9E3779B97F4A7C15ull * (blockIdx.x * blockDim.x + threadIdx.x);
    unsigned long long nbad_div = 0, nbad_sqrt = 0, nbad_div2 = 0, nslow = 0;
    for (int it = 0; it < iters; ++it) {
        double a = rnd_double(s, 300), b = fabs(rnd_double(s, 300));
        if ((it & 3) == 1) {            // near-boundary quotient: a = RN(q * b) +- few ulps
            const double q = rnd_double(s, 40);
            b = fabs(rnd_double(s, 40));
            a = q * b;
            const int k = (int)(splitmix(s) % 5) - 2;
            a = __longlong_as_double(__double_as_longlong(a) + k);
        } else if ((it & 3) == 2) {     // the sweep's own shapes: rr = sqrt(f^2 + g^2), f / rr, g / rr
            const double f = fabs(rnd_double(s, 8)), g = rnd_double(s, 8);
            b = sqrt(fma(f, f, g * g));
            a = (it & 4) ? f : g;
        }
        double a2 = rnd_double(s, 200);
        // single and shared-denominator division
        if (amhd::fast_div_ok(a, b) && amhd::fast_div_ok(a2, b)) {
            double q1, q2;
            amhd::div2_same_den(a, a2, b, q1, q2);
            if (__double_as_longlong(q1) != __double_as_longlong(a / b)) ++nbad_div;
            if (__double_as_longlong(q2) != __double_as_longlong(a2 / b)) ++nbad_div2;
        } else ++nslow;
        const double x = fabs(a);
        if (amhd::fast_sqrt_ok(x)) {
            if (__double_as_longlong(amhd::sqrt_fast(x)) != __double_as_longlong(sqrt(x))) ++nbad_sqrt;
        }
    }
    atomicAdd(bad + 0, nbad_div);
    atomicAdd(bad + 1, nbad_div2);
    atomicAdd(bad + 2, nbad_sqrt);
    atomicAdd(bad + 3, nslow);
}

int main() {
    unsigned long long* bad;
    CK(cudaMallocManaged(&bad, 4 * sizeof(unsigned long long)));
    for (int i = 0; i < 4; ++i) bad[i] = 0;
    const int blocks = 148 * 8, threads = 256, iters = 20000;
    probe<<<blocks, threads>>>(bad, iters);
    CK(cudaDeviceSynchronize());
    const double total = (double)blocks * threads * iters;
    printf("cases %.3g: div mismatches %llu, shared-denominator second quotient mismatches %llu, sqrt mismatches %llu, guarded (slow path) %llu\n",
           total, bad[0], bad[1], bad[2], bad[3]);
    return (bad[0] || bad[1] || bad[2]) ? 1 : 0;
}
