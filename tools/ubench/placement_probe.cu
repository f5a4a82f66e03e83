// placement_probe.cu -- how does the block scheduler place a 1024-CTA grid whose occupancy limit is 7 CTAs/SM
// (the K1T16 launch: 128 threads, 72 registers, 26.6 KB dynamic shared memory) on a B200's 148 SMs?
// Every CTA records its SM id and start / end times while spinning for a fixed number of cycles.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o placement_probe placement_probe.cu
#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <map>
#include <vector>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

__global__ void __launch_bounds__(128, 7) probe(unsigned* smid, unsigned long long* t0, unsigned long long* t1, long long spin, double* sink) {
    extern __shared__ double sm[];
    unsigned long long a, b;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(a));
    unsigned id;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(id));
    sm[threadIdx.x] = threadIdx.x;
    const long long c0 = clock64();
    double acc = 0;
    while (clock64() - c0 < spin) acc += sm[(threadIdx.x * 7) & 127];
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(b));
    if (threadIdx.x == 0) { smid[blockIdx.x] = id; t0[blockIdx.x] = a; t1[blockIdx.x] = b; }
    if (acc == 1.2345) sink[0] = acc;
}

int main(int argc, char** argv) {
    const int grid = argc > 1 ? atoi(argv[1]) : 1024;
    const size_t smem = 26624;
    unsigned* smid; unsigned long long *t0, *t1; double* sink;
    CK(cudaMallocManaged(&smid, grid * 4)); CK(cudaMallocManaged(&t0, grid * 8)); CK(cudaMallocManaged(&t1, grid * 8)); CK(cudaMalloc(&sink, 8));
    CK(cudaFuncSetAttribute(probe, cudaFuncAttributePreferredSharedMemoryCarveout, 84));
    int per_sm = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, probe, 128, smem));
    probe<<<grid, 128, smem>>>(smid, t0, t1, 1000, sink);
    CK(cudaDeviceSynchronize());
    probe<<<grid, 128, smem>>>(smid, t0, t1, 2000000, sink);      // ~1 ms
    CK(cudaDeviceSynchronize());
    std::map<unsigned, int> cnt;
    unsigned long long tmin = ~0ull, tmax = 0;
    for (int i = 0; i < grid; ++i) { cnt[smid[i]]++; tmin = std::min(tmin, t0[i]); tmax = std::max(tmax, t1[i]); }
    std::map<int, int> hist;
    for (auto& kv : cnt) hist[kv.second]++;
    int late = 0;
    for (int i = 0; i < grid; ++i) if (t0[i] - tmin > 200000) ++late;     // started > 0.2 ms after the first CTA
    printf("grid %d, occupancy API %d CTAs/SM, SMs used %zu, kernel span %.3f ms, CTAs that started late (second wave) %d\n", grid, per_sm,
           cnt.size(), (tmax - tmin) * 1e-6, late);
    for (auto& kv : hist) printf("  %d SMs hold %d CTAs\n", kv.second, kv.first);
    return 0;
}
