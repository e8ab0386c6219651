// Stand-alone check of the tcgen05 building blocks (descriptors, swizzled operand layout, TMEM
// load path) used by the rollout engine:  out[n][m] = sum_k W[m][k] * X[n][k]  with the weights
// as the 128-row A operand and the (few) tokens as the B operand.  Debug entry point only.
#include "umma.cuh"
#include "ro_kernel.h"

namespace sfb {

__global__ void __launch_bounds__(160, 1) umma_test_kernel(const __half* __restrict__ Wp, const float* __restrict__ X,
                                                            float* __restrict__ out, int M, int N, int K) {
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char* wtile = smem;                       // 16 KB
    unsigned char* xact = smem + 16384;                // [K/64][N][128 B]
    const int kpt = K >> 6;
    uint64_t* bars = reinterpret_cast<uint64_t*>(xact + (size_t)kpt * N * 128);
    uint64_t* full = bars;
    uint64_t* mmadone = bars + 1;
    uint64_t* accfull = bars + 2;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 3);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (tid == 0) { mbar_init(full, 1); mbar_init(mmadone, 1); mbar_init(accfull, 1); fence_mbar_init(); }
    if (warp == 4) tmem_alloc(tmem_ptr, 128);
    for (int i = tid; i < N * K; i += blockDim.x) {
        const int r = i / K, k = i % K;
        *reinterpret_cast<__half*>(xact + swz_off(r, k, N * 128)) = __float2half_rn(X[i]);
    }
    fence_proxy_async();
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem = *tmem_ptr;
    const uint32_t idesc = umma_idesc_f16(128, N);

    const int tiles = M >> 7;
    uint32_t nfull = 0, ndone = 0;
    for (int t = 0; t < tiles; ++t) {
        if (warp == 4) {
            if (lane == 0) {
                const uint64_t pol = l2_policy_evict_last();
                for (int kb = 0; kb < kpt; ++kb) {
                    mbar_arrive_expect_tx(full, 16384);
                    bulk_g2s(wtile, Wp + ((size_t)t * kpt + kb) * 8192, 16384, full, pol);
                    mbar_wait(full, nfull & 1); ++nfull;
                    tcgen05_fence_after();
#pragma unroll
                    for (int k4 = 0; k4 < 4; ++k4) {
                        const uint64_t da = umma_smem_desc(smem_u32(wtile) + k4 * 32);
                        const uint64_t db = umma_smem_desc(smem_u32(xact) + kb * N * 128 + k4 * 32);
                        umma_f16(tmem, da, db, idesc, (kb | k4) != 0);
                    }
                    umma_commit(mmadone);
                    mbar_wait(mmadone, ndone & 1); ++ndone;
                }
                mbar_arrive(accfull);
            }
            __syncwarp();
        } else {
            mbar_wait(accfull, t & 1);
            tcgen05_fence_after();
            const int f = 32 * warp + lane;
            for (int c0 = 0; c0 < N; c0 += 8) {
                float v[8];
                tmem_ld8(tmem + ((uint32_t)(32 * warp) << 16) + c0, v);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 8; ++i) out[(size_t)(c0 + i) * M + t * 128 + f] = v[i];
            }
            tcgen05_fence_before();
        }
        __syncthreads();
    }
    if (warp == 4) tmem_dealloc(tmem, 128);
}

cudaError_t umma_test_launch(const __half* Wp, const float* X, float* out, int M, int N, int K, cudaStream_t st) {
    const size_t smem = 16384 + (size_t)(K >> 6) * N * 128 + 64;
    cudaError_t e = cudaFuncSetAttribute(umma_test_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    umma_test_kernel<<<1, 160, smem, st>>>(Wp, X, out, M, N, K);
    return cudaGetLastError();
}

}  // namespace sfb
