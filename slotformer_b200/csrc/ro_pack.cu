// Weight packing for the rollout engines: fp32 state_dict matrices -> fp16 operand tiles (sfb_rollout_prepare).
#include "common.cuh"
#include "ro_kernel.h"

namespace sfb {

// fp32 [Nw][Kd] -> fp16 128x64 weight tiles (two 64x64 128B-swizzled panels each), tile-major
// (tile = 128 output features), kb-minor; rows beyond Nw are zero.
// lo: the residual fp16(w - fp16(w)) instead of fp16(w).
__global__ void ro_pack2_kernel(const float* __restrict__ src, __half* __restrict__ dst, int Nw, int Kd, bool lo) {
    const int npad = (Nw + 127) & ~127;
    const size_t total = (size_t)npad * Kd;
    const int kpt = Kd >> 6;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int n = (int)(i / Kd), k = (int)(i % Kd);
        const int nb = n >> 6, r = n & 63, kb = k >> 6, kk = k & 63;
        const size_t off = (((size_t)(nb >> 1) * kpt + kb) * 2 + (nb & 1)) * 4096 + r * 64 +
                           ((((kk >> 3) ^ (r & 7)) << 3) | (kk & 7));
        const float w = n < Nw ? src[(size_t)n * Kd + k] : 0.f;
        const __half hi = __float2half_rn(w);
        dst[off] = lo ? __float2half_rn(w - __half2float(hi)) : hi;
    }
}

cudaError_t ro_pack2_launch(const float* src, __half* dst, int N, int Kd, cudaStream_t st, bool lo) {
    const size_t n = (size_t)((N + 127) & ~127) * Kd;
    int blocks = (int)((n + 255) / 256);
    if (blocks > 1184) blocks = 1184;
    ro_pack2_kernel<<<blocks, 256, 0, st>>>(src, dst, N, Kd, lo);
    return cudaGetLastError();
}

}  // namespace sfb
