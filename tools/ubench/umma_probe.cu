// umma_probe.cu -- stand-alone check of the tcgen05 building blocks the split-bf16 MALA kernel (K3T) is made of:
//   (1) hand-written K-major SWIZZLE_128B operand tiles in shared memory (the layout TMA produces and UMMA consumes),
//   (2) shared-memory / instruction descriptors, tcgen05.mma kind::f16 (bf16 x bf16 -> f32 in TMEM), tcgen05.commit,
//   (3) tcgen05.ld 32x32b of the accumulator by the warps that own the TMEM lanes,
//   (4) the same operands brought in by TMA (cp.async.bulk.tensor.2d, 128B swizzle) instead of by threads.
// D[128 x N] = A[128 x K] * B[N x K]^T with K = 128 (two 64-element K blocks, 8 MMAs of K = 16), compared with a CPU
// reference on the bf16-rounded inputs.   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o umma_probe umma_probe.cu -lcuda
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cmath>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* b, int c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c)); }
__device__ __forceinline__ void mbar_wait(unsigned long long* b, unsigned ph) {
    asm volatile("{\n.reg .pred p;\nW:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D;\nbra W;\nD:\n}\n" ::"r"(smem_u32(b)), "r"(ph) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* b, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_2d(void* dst, const CUtensorMap* map, int c0, int c1, unsigned long long* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar)) : "memory");
}

// K-major SWIZZLE_128B descriptor: rows of 128 bytes, 8-row groups 1024 bytes apart
__device__ __forceinline__ unsigned long long make_desc(const void* p) {
    unsigned long long d = 0;
    d |= (unsigned long long)((smem_u32(p) >> 4) & 0x3FFF);
    d |= (unsigned long long)1 << 16;            // leading byte offset (unused with swizzle) = 1
    d |= (unsigned long long)64 << 32;           // stride byte offset = 1024 B >> 4
    d |= (unsigned long long)1 << 46;            // descriptor version 1 (Blackwell)
    d |= (unsigned long long)2 << 61;            // SWIZZLE_128B
    return d;
}
__host__ __device__ constexpr unsigned make_idesc(int M, int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((unsigned)(N >> 3) << 17) | ((unsigned)(M >> 4) << 24);   // f32 accum, bf16 x bf16, K-major both
}
__device__ __forceinline__ void umma(unsigned tmem_d, unsigned long long da, unsigned long long db, unsigned idesc, unsigned acc) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n"
                 ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_commit(unsigned long long* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// byte offset of element (row, k) of a [rows x 64] bf16 K-block tile in the 128B-swizzled K-major layout
__host__ __device__ inline int sw128_off(int row, int k) {
    const int chunk = k >> 3;
    return row * 128 + ((chunk ^ (row & 7)) << 4) + (k & 7) * 2;
}

template <int N, bool USE_TMA>
__global__ void __launch_bounds__(192) probe_kernel(const __nv_bfloat16* A, const __nv_bfloat16* B, float* D,
                                                    const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB) {
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char* sA = smem;                       // 2 K blocks x [128 x 128 B]
    unsigned char* sB = smem + 2 * 128 * 128;       // 2 K blocks x [N x 128 B]
    unsigned long long* bars = (unsigned long long*)(sB + 2 * N * 128);
    unsigned* tmem_ptr = (unsigned*)(bars + 4);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        mbar_init(bars + 0, 1);                     // MMA done
        mbar_init(bars + 1, 1);                     // TMA full
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "r"(N < 32 ? 32 : N));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (!USE_TMA) {
        // every thread writes elements with the swizzle applied by hand
        for (int idx = threadIdx.x; idx < 128 * 128; idx += blockDim.x) {
            const int row = idx / 128, k = idx % 128;
            *(__nv_bfloat16*)(sA + (k >> 6) * 128 * 128 + sw128_off(row, k & 63)) = A[idx];
        }
        for (int idx = threadIdx.x; idx < N * 128; idx += blockDim.x) {
            const int row = idx / 128, k = idx % 128;
            *(__nv_bfloat16*)(sB + (k >> 6) * N * 128 + sw128_off(row, k & 63)) = B[idx];
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy writes -> visible to the tensor core (async proxy)
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const unsigned tmem = *tmem_ptr;
    if (USE_TMA && warp == 0 && lane == 0) {
        mbar_expect_tx(bars + 1, 2 * 128 * 128 + 2 * N * 128);
        for (int kb = 0; kb < 2; ++kb) {
            tma_2d(sA + kb * 128 * 128, &mapA, kb * 64, 0, bars + 1);
            tma_2d(sB + kb * N * 128, &mapB, kb * 64, 0, bars + 1);
        }
    }
    if (warp == 1 && lane == 0) {
        if (USE_TMA) { mbar_wait(bars + 1, 0); asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
        constexpr unsigned idesc = make_idesc(128, N);
        for (int kb = 0; kb < 2; ++kb)
            for (int kk = 0; kk < 4; ++kk) {
                const unsigned long long da = make_desc(sA + kb * 128 * 128 + kk * 32);
                const unsigned long long db = make_desc(sB + kb * N * 128 + kk * 32);
                umma(tmem, da, db, idesc, (kb | kk) ? 1u : 0u);
            }
        umma_commit(bars + 0);
    }
    if (warp >= 2) {
        mbar_wait(bars + 0, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int lg = warp & 3;                                   // TMEM lane group this warp may access
        const int row = 32 * lg + lane;
        for (int c0 = 0; c0 < N; c0 += 16) {
            unsigned r[16];
            const unsigned taddr = tmem + ((unsigned)(32 * lg) << 16) + (unsigned)c0;
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                         : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                           "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                         : "r"(taddr));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            for (int i = 0; i < 16; ++i) D[row * N + c0 + i] = __uint_as_float(r[i]);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(N < 32 ? 32 : N));
}

static CUtensorMap make_map(const void* base, int rows, int cols /* K, inner */) {
    CUtensorMap m;
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)cols * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = cuTensorMapEncodeTiled(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void*)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("cuTensorMapEncodeTiled failed: %d\n", (int)r); exit(1); }
    return m;
}

template <int N, bool USE_TMA>
static int run() {
    std::vector<__nv_bfloat16> hA(128 * 128), hB(N * 128);
    std::vector<float> fA(128 * 128), fB(N * 128);
    unsigned s = 12345u;
    auto rnd = [&]() { s = s * 1664525u + 1013904223u; return ((s >> 8) & 0xFFFF) / 65536.0f - 0.5f; };
    for (size_t i = 0; i < hA.size(); ++i) { hA[i] = __float2bfloat16(rnd()); fA[i] = __bfloat162float(hA[i]); }
    for (size_t i = 0; i < hB.size(); ++i) { hB[i] = __float2bfloat16(rnd()); fB[i] = __bfloat162float(hB[i]); }
    __nv_bfloat16 *dA, *dB; float* dD;
    CK(cudaMalloc(&dA, hA.size() * 2)); CK(cudaMalloc(&dB, hB.size() * 2)); CK(cudaMalloc(&dD, 128 * N * 4));
    CK(cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemset(dD, 0xff, 128 * N * 4));
    const CUtensorMap mA = make_map(dA, 128, 128), mB = make_map(dB, N, 128);
    const size_t smem = 2 * 128 * 128 + 2 * N * 128 + 64 + 1024;
    CK(cudaFuncSetAttribute(probe_kernel<N, USE_TMA>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    probe_kernel<N, USE_TMA><<<1, 192, smem>>>(dA, dB, dD, mA, mB);
    CK(cudaDeviceSynchronize());
    std::vector<float> hD(128 * N);
    CK(cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost));
    double maxerr = 0; int bad = 0;
    for (int m = 0; m < 128; ++m)
        for (int n = 0; n < N; ++n) {
            double ref = 0;
            for (int k = 0; k < 128; ++k) ref += (double)fA[m * 128 + k] * fB[n * 128 + k];
            const double err = fabs(ref - hD[m * N + n]);
            if (!(err <= 1e-3)) ++bad;
            if (err > maxerr || err != err) maxerr = err;
        }
    printf("N=%3d %s: max |D - ref| = %.3g, %d of %d entries off\n", N, USE_TMA ? "TMA   " : "manual", maxerr, bad, 128 * N);
    cudaFree(dA); cudaFree(dB); cudaFree(dD);
    return bad;
}

int main() {
    CK(cudaFree(0));
    int bad = 0;
    bad += run<64, false>();
    bad += run<128, false>();
    bad += run<64, true>();
    bad += run<128, true>();
    printf(bad ? "PROBE FAILED\n" : "PROBE OK\n");
    return bad ? 1 : 0;
}
