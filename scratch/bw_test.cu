// Per-SM L2->smem streaming bandwidth: cp.async.bulk (1-D) vs cp.async.bulk.tensor (2-D tensor map) vs LDG.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(s32(b)), "r"(c)); }
__device__ __forceinline__ void mbar_expect(uint64_t* b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(s32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t par) {
    uint32_t ok = 0;
    while (!ok) asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0,1,0,p;\n}" : "=r"(ok) : "r"(s32(b)), "r"(par) : "memory");
}
__device__ __forceinline__ void bulk1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" :: "r"(s32(dst)), "l"(src), "r"(bytes), "r"(s32(bar)) : "memory");
}
__device__ __forceinline__ void tma2d(void* dst, const CUtensorMap* map, int x, int y, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 :: "r"(s32(dst)), "l"(map), "r"(s32(bar)), "r"(x), "r"(y) : "memory");
}
// mode 0: one thread, `chunk`-byte 1-D bulk copies, NST stages of 16 KB each (stage = 16384/chunk copies)
// mode 1: tensor-map 2-D copies of 16 KB (box 64 fp16 x 128 rows)
// mode 2: 256 threads LDG.128 -> STS (no async)
// mode 3: `nprod` producer threads (one per warp), each owning stages s % nprod
__global__ void __launch_bounds__(288, 1) bw_kernel(const unsigned char* src, size_t total_bytes, int iters, int mode, int chunk, int nst,
                                                     const __grid_constant__ CUtensorMap map, unsigned long long* out_cycles, float* sink) {
    extern __shared__ __align__(1024) unsigned char smem[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + nst * 16384);
    const int tid = threadIdx.x;
    if (tid == 0) { for (int s = 0; s < nst; ++s) mbar_init(&full[s], 1); asm volatile("fence.mbarrier_init.release.cluster;"); }
    __syncthreads();
    const int nstages_total = (int)(total_bytes / 16384) * iters;
    const int per_iter = (int)(total_bytes / 16384);
    unsigned long long t0 = clock64();
    float acc = 0.f;
    if (mode == 2) {
        for (int n = 0; n < nstages_total; ++n) {
            const uint4* g = reinterpret_cast<const uint4*>(src + (size_t)(n % per_iter) * 16384);
            uint4* d = reinterpret_cast<uint4*>(smem + (n % nst) * 16384);
            for (int i = tid; i < 1024; i += 288) d[i] = __ldg(g + i);
            __syncthreads();
        }
    } else {
        // consumer = warp 0 lane 0 waits in order; producer(s) issue up to nst ahead
        if (tid == 32) {   // producer thread
            int issued = 0;
            // simple scheme: issue nst stages, then refill one each time consumer signals via a shared counter
            volatile int* done = reinterpret_cast<volatile int*>(smem + nst * 16384 + 256);
            while (issued < nstages_total) {
                while (issued - *done >= nst) { }
                const int s = issued % nst;
                mbar_expect(&full[s], 16384);
                const unsigned char* g = src + (size_t)(issued % per_iter) * 16384;
                if (mode == 0) { for (int c = 0; c < 16384; c += chunk) bulk1d(smem + s * 16384 + c, g + c, chunk, &full[s]); }
                else { tma2d(smem + s * 16384, &map, 0, (issued % per_iter) * 128, &full[s]); }
                ++issued;
            }
        } else if (tid == 0) {
            volatile int* done = reinterpret_cast<volatile int*>(smem + nst * 16384 + 256);
            *done = 0;
            for (int n = 0; n < nstages_total; ++n) {
                const int s = n % nst;
                mbar_wait(&full[s], (n / nst) & 1);
                acc += reinterpret_cast<float*>(smem + s * 16384)[n & 1023];
                *done = n + 1;
            }
        }
    }
    __syncthreads();
    unsigned long long t1 = clock64();
    if (tid == 0) { out_cycles[blockIdx.x] = t1 - t0; sink[blockIdx.x] = acc; }
}
typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main() {
    const size_t total = 1664 * 1024;   // 1.66 MB of weights, L2 resident
    unsigned char* src; CK(cudaMalloc(&src, total)); CK(cudaMemset(src, 1, total));
    unsigned long long* cyc; CK(cudaMalloc(&cyc, 148 * 8)); float* sink; CK(cudaMalloc(&sink, 148 * 4));
    void* fn = nullptr; cudaDriverEntryPointQueryResult qres;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    CUtensorMap map;
    cuuint64_t dims[2] = {64, total / 128}; cuuint64_t strides[1] = {128}; cuuint32_t box[2] = {64, 128}; cuuint32_t estr[2] = {1, 1};
    CUresult r = ((EncodeFn)fn)(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, src, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode result %d\n", (int)r);
    CK(cudaFuncSetAttribute(bw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 16384 + 1024));
    int clk; CK(cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0));
    struct { int mode, chunk, nst; const char* name; } cfgs[] = {
        {0, 16384, 8, "1-D bulk 16 KB x8 stages"}, {0, 4096, 8, "1-D bulk 4 KB x8"}, {0, 1024, 8, "1-D bulk 1 KB x8"},
        {0, 16384, 4, "1-D bulk 16 KB x4"}, {1, 0, 8, "tensor 2-D 16 KB x8"}, {1, 0, 4, "tensor 2-D 16 KB x4"}, {2, 0, 8, "LDG.128 256 thr"}};
    for (int grid : {1, 64, 148}) for (auto& c : cfgs) {
        const int iters = 20;
        bw_kernel<<<grid, 288, c.nst * 16384 + 1024>>>(src, total, 2, c.mode, c.chunk, c.nst, map, cyc, sink); CK(cudaDeviceSynchronize());
        cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b); cudaEventRecord(a);
        bw_kernel<<<grid, 288, c.nst * 16384 + 1024>>>(src, total, iters, c.mode, c.chunk, c.nst, map, cyc, sink);
        cudaEventRecord(b); CK(cudaDeviceSynchronize()); float ms; cudaEventElapsedTime(&ms, a, b);
        printf("grid %3d  %-28s %7.1f GB/s per SM  (%.1f us)\n", grid, c.name, total * iters / (ms * 1e-3) / 1e9, ms * 1e3);
    }
    return 0;
}
