// Encoder tail fused up to Slot Attention's own LayerNorm, C = 128.
//
// Reference (base_slots/models/savi.py:367-377, utils.py:52-63), per pixel of the 64-channel CNN output:
//     x   = cnn_out[:, p] + dense(grid[p])                 SoftPositionEmbed
//     y   = W2 relu(W1 LN_64(x) + b1) + b2                 encoder_out_layer
// followed by SlotAttention.norm_inputs (savi.py:66).  The reference materialises y as the fp32 [frames, 4096, 128]
// feature grid (2.1 MB per frame, THE input stream of Slot Attention).  Here one kernel reads the 64-channel CNN
// output once (1.05 MB per frame) and writes t = (y - mean) * rstd directly as the fp16 operand tiles the tcgen05
// Slot Attention passes consume (the x^ ring of sa_pass_tc.cu): the fp32 feature grid never exists, and every
// Slot Attention iteration -- the first one included -- is a ring pass.
//
//   producer warp   tensor-map TMA of [64 ch x 128 px] fp32 boxes (NCHW input: pixels contiguous); the packed
//                   weights once per CTA
//   4 LN64 warps    thread = pixel: + positional embedding (a_c y + b_c x + d_c, the 4->64 dense layer collapsed
//                   to 3 constants per channel), LayerNorm over the 64 channels in registers, fp16 operand row
//   MMA warp        D1[128 px x 128] = U x W1'^T (LN affine folded into W1', b1'), D2 = H x W2^T; accumulators in
//                   TMEM, double-buffered (512 columns)
//   4 epi1 warps    thread = pixel: relu(D1 + b1') -> fp16 operand rows of the second GEMM
//   4 epi2 warps    thread = pixel: y = D2 + b2, LayerNorm over the 128 features in registers, t -> fp16 tile
//                   (written over the consumed H operand), TMA bulk store of the tile to the ring
// Operands are fp16 (single term): measured against the reference this perturbs the extracted slots by 1e-4
// relative (tests/test_encoder_tail.py), a tenth of the stated tolerance.
#include "umma.cuh"
#include "sa_kernel.h"

#include <cuda.h>

namespace sfb {

namespace {

constexpr float ET_LN_EPS = 1e-5f;
constexpr int ET_CIN = 64, ET_C = 128, ET_TILE_PX = 128;
constexpr int ET_IN_BYTES = ET_CIN * ET_TILE_PX * 4;       // 32 KB: [64 ch][128 px] fp32
constexpr int ET_A1_BYTES = ET_TILE_PX * 128;              // 16 KB: [128 px][64 ch] fp16, one swizzled panel
constexpr int ET_PANEL = ET_TILE_PX * 128;                 // 16 KB
constexpr int ET_TILE_BYTES = 2 * ET_PANEL;                // 32 KB: [128 px][128 ch] fp16, two panels
constexpr int ET_W1_BYTES = ET_C * 128;                    // 16 KB: [128 out][64 in]
constexpr int ET_W2_BYTES = 2 * ET_C * 128;                // 32 KB: [128 out][128 in], two k panels

constexpr int OFF_T = 0;                                   // 2 x 32 KB  H operand / t tile
constexpr int OFF_A1 = OFF_T + 2 * ET_TILE_BYTES;          // 2 x 16 KB
constexpr int OFF_W1 = OFF_A1 + 2 * ET_A1_BYTES;           // 16 KB
constexpr int OFF_W2 = OFF_W1 + ET_W1_BYTES;               // 32 KB
constexpr int OFF_IN = OFF_W2 + ET_W2_BYTES;               // 2 x 32 KB raw stages
constexpr int OFF_PAR = OFF_IN + 2 * ET_IN_BYTES;          // b1'[128] b2[128] pos[64][4]
constexpr int ET_PAR_FLOATS = 128 + 128 + 64 * 4;
constexpr int OFF_BARS = OFF_PAR + ET_PAR_FLOATS * 4;
enum { EB_IN_FULL = 0, EB_IN_EMPTY = 2, EB_A1_READY = 4, EB_A1_FREE = 6, EB_D1_READY = 8, EB_A2_READY = 10,
       EB_D2_READY = 12, EB_D2_FREE = 14, EB_T_FREE = 16, EB_W_READY = 18, EB_WORDS = 20 };
constexpr int OFF_TMEM = OFF_BARS + EB_WORDS * 8;
constexpr int ET_SMEM = OFF_TMEM + 16;
static_assert(OFF_A1 % 1024 == 0 && OFF_W1 % 1024 == 0 && OFF_W2 % 1024 == 0 && OFF_IN % 1024 == 0, "alignment");
static_assert(ET_SMEM <= 232448, "encoder tail: shared memory budget");

typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float lo, float hi) {
    f32x2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r;
}
__device__ __forceinline__ float lo2(f32x2 v) { float a, b; asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); return a; }
__device__ __forceinline__ float hi2(f32x2 v) { float a, b; asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); return b; }

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                 "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
                 "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                   "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                   "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                 : "r"(taddr));
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ void tma_load_3d(void* dst_smem, const CUtensorMap* tmap, int c0, int c1, int c2,
                                            uint64_t* bar, uint64_t policy) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
                 " [%0], [%1, {%2, %3, %4}], [%5], %6;"
                 :: "r"(smem_u32(dst_smem)), "l"(tmap), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar)), "l"(policy)
                 : "memory");
}

__device__ __forceinline__ void mbar_arrive_after(uint64_t* bar, float dep) {
    uint32_t z;
    asm volatile("and.b32 %0, %1, 0;" : "=r"(z) : "r"(__float_as_uint(dep)));
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(bar) + z) : "memory");
}

__device__ __forceinline__ void mbar_wait_lane0(uint64_t* bar, uint32_t parity, int lane) {
    if (lane == 0) { while (!mbar_try_wait(bar, parity)) { } }
    __syncwarp();
}

}  // namespace

struct EncTailParams {
    __half* tiles;             // [frames][tiles_frame][32 KB]
    const __half* w1;          // packed operand images (enc_tail_prep_kernel)
    const __half* w2;
    const float* par;          // b1'[128] | b2[128] | pos[64][4] = (a, b, d, 0)
    int frames, N, W, tiles_frame;
    float inv_hm1, inv_wm1;    // 1 / (H - 1), 1 / (W - 1)
};

// NHWC: the CNN output is channels-last ([frames][N pixels][64 channels], what cuDNN's tensor-core convolutions produce
// natively): the raw stage is two 128-byte-swizzled panels [128 px][32 ch] fp32 instead of one [64 ch][128 px] slab
template <bool NHWC>
__global__ void __launch_bounds__(512, 1) enc_tail_kernel(const EncTailParams p, const __grid_constant__ CUtensorMap tmap) {
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char* tbuf = smem + OFF_T;
    unsigned char* a1buf = smem + OFF_A1;
    unsigned char* w1s = smem + OFF_W1;
    unsigned char* w2s = smem + OFF_W2;
    unsigned char* inbuf = smem + OFF_IN;
    float* par = reinterpret_cast<float*>(smem + OFF_PAR);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BARS);
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + OFF_TMEM);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (tid == 0) {
        for (int s = 0; s < 2; ++s) {
            mbar_init(&bars[EB_IN_FULL + s], 1);   mbar_init(&bars[EB_IN_EMPTY + s], 4);
            mbar_init(&bars[EB_A1_READY + s], 4);  mbar_init(&bars[EB_A1_FREE + s], 1);
            mbar_init(&bars[EB_D1_READY + s], 1);  mbar_init(&bars[EB_A2_READY + s], 4);
            mbar_init(&bars[EB_D2_READY + s], 1);  mbar_init(&bars[EB_D2_FREE + s], 4);
            mbar_init(&bars[EB_T_FREE + s], 4);
        }
        mbar_init(&bars[EB_W_READY], 1);
        fence_mbar_init();
    }
    for (int i = tid; i < ET_PAR_FLOATS; i += 512) par[i] = __ldg(p.par + i);
    if (warp == 13) tmem_alloc(tmem_ptr, 512);
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem = *tmem_ptr;
    const int total_tiles = p.frames * p.tiles_frame;
    const int my_tiles = (total_tiles > (int)blockIdx.x) ? (total_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
    const int row = 32 * (warp & 3) + lane;                 // pixel of the tile / TMEM lane
    const uint32_t lane_base = tmem + ((uint32_t)(32 * (warp & 3)) << 16);
    uint32_t xo[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) xo[j] = (uint32_t)((j ^ (row & 7)) << 4);

    if (warp < 4) {
        // ------------------------------- LN64 warps -------------------------------
        asm volatile("setmaxnreg.dec.sync.aligned.u32 104;");
        const float* pos = par + 256;
        for (int i = 0; i < my_tiles; ++i) {
            const int g = (int)blockIdx.x + i * (int)gridDim.x;
            const int tif = g % p.tiles_frame;
            const int b = i & 1;
            const int px = tif * ET_TILE_PX + row;
            const float yy = (float)(px / p.W) * p.inv_hm1, xx = (float)(px % p.W) * p.inv_wm1;
            const bool valid = px < p.N;                     // pixels beyond N: zero rows (a fully out-of-range tile is not even loaded)
            mbar_wait_lane0(&bars[EB_IN_FULL + b], (i >> 1) & 1, lane);
            const float* src = reinterpret_cast<const float*>(inbuf + (size_t)b * ET_IN_BYTES) + row;
            float v[ET_CIN];
            float s1 = 0.f, s2 = 0.f;
            if (NHWC) {
                const unsigned char* prow = inbuf + (size_t)b * ET_IN_BYTES + row * 128;
#pragma unroll
                for (int q = 0; q < 2; ++q)
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float4 w = *reinterpret_cast<const float4*>(prow + q * (ET_IN_BYTES / 2) + xo[j]);
                        v[32 * q + 4 * j] = w.x; v[32 * q + 4 * j + 1] = w.y; v[32 * q + 4 * j + 2] = w.z; v[32 * q + 4 * j + 3] = w.w;
                    }
            }
#pragma unroll
            for (int c = 0; c < ET_CIN; ++c) {
                const float4 pc = *reinterpret_cast<const float4*>(pos + 4 * c);
                const float raw = NHWC ? v[c] : src[c * ET_TILE_PX];
                v[c] = valid ? raw + fmaf(pc.x, yy, fmaf(pc.y, xx, pc.z)) : 0.f;
                s1 += v[c];
                s2 = fmaf(v[c], v[c], s2);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive_after(&bars[EB_IN_EMPTY + b], s1);
            const float mu = s1 * (1.f / ET_CIN);
            const float var = fmaxf(fmaf(-mu, mu, s2 * (1.f / ET_CIN)), 0.f);
            const float rstd = rsqrtf(var + ET_LN_EPS);
            const float nb = -mu * rstd;
            mbar_wait_lane0(&bars[EB_A1_FREE + b], ((i >> 1) & 1) ^ 1, lane);
            unsigned char* arow = a1buf + (size_t)b * ET_A1_BYTES + row * 128;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                uint4 pk;
                pk.x = pack_h2(fmaf(v[8 * j], rstd, nb), fmaf(v[8 * j + 1], rstd, nb));
                pk.y = pack_h2(fmaf(v[8 * j + 2], rstd, nb), fmaf(v[8 * j + 3], rstd, nb));
                pk.z = pack_h2(fmaf(v[8 * j + 4], rstd, nb), fmaf(v[8 * j + 5], rstd, nb));
                pk.w = pack_h2(fmaf(v[8 * j + 6], rstd, nb), fmaf(v[8 * j + 7], rstd, nb));
                *reinterpret_cast<uint4*>(arow + xo[j]) = pk;
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bars[EB_A1_READY + b]);
        }
    } else if (warp < 8) {
        // ------------------------------- epilogue 1: relu(D1 + b1') -> H operand -------------------------------
        asm volatile("setmaxnreg.dec.sync.aligned.u32 96;");
        const float* b1 = par;
        for (int i = 0; i < my_tiles; ++i) {
            const int b = i & 1;
            mbar_wait_lane0(&bars[EB_D1_READY + b], (i >> 1) & 1, lane);
            tcgen05_fence_after();
            mbar_wait_lane0(&bars[EB_T_FREE + b], ((i >> 1) & 1) ^ 1, lane);     // the tile of i-2 has left this buffer
            unsigned char* hrow = tbuf + (size_t)b * ET_TILE_BYTES + row * 128;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                float d[32];
                tmem_ld32(lane_base + (uint32_t)(128 * b + 32 * q), d);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int c0 = 32 * q + 8 * j;
                    uint4 pk;
                    pk.x = pack_h2(fmaxf(d[8 * j] + b1[c0], 0.f), fmaxf(d[8 * j + 1] + b1[c0 + 1], 0.f));
                    pk.y = pack_h2(fmaxf(d[8 * j + 2] + b1[c0 + 2], 0.f), fmaxf(d[8 * j + 3] + b1[c0 + 3], 0.f));
                    pk.z = pack_h2(fmaxf(d[8 * j + 4] + b1[c0 + 4], 0.f), fmaxf(d[8 * j + 5] + b1[c0 + 5], 0.f));
                    pk.w = pack_h2(fmaxf(d[8 * j + 6] + b1[c0 + 6], 0.f), fmaxf(d[8 * j + 7] + b1[c0 + 7], 0.f));
                    *reinterpret_cast<uint4*>(hrow + (q >> 1) * ET_PANEL + xo[(4 * q + j) & 7]) = pk;
                }
            }
            fence_proxy_async();
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bars[EB_A2_READY + b]);
        }
    } else if (warp < 12) {
        // ------------------------------- epilogue 2: y = D2 + b2, LayerNorm, t tile, store -------------------------------
        asm volatile("setmaxnreg.inc.sync.aligned.u32 200;");
        const float* b2 = par + 128;
        const uint64_t pol = l2_policy_evict_last();
        const int sub = warp & 3;
        for (int i = 0; i < my_tiles; ++i) {
            const int g = (int)blockIdx.x + i * (int)gridDim.x;
            const int f = g / p.tiles_frame, tif = g % p.tiles_frame;
            const int b = i & 1;
            const int px = tif * ET_TILE_PX + row;
            mbar_wait_lane0(&bars[EB_D2_READY + b], (i >> 1) & 1, lane);
            tcgen05_fence_after();
            float y[ET_C];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                float d[32];
                tmem_ld32(lane_base + (uint32_t)(256 + 128 * b + 32 * q), d);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 32; ++j) y[32 * q + j] = d[j] + b2[32 * q + j];
            }
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bars[EB_D2_FREE + b]);
            float s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int c = 0; c < ET_C; ++c) { s1[c & 3] += y[c]; s2[c & 3] = fmaf(y[c], y[c], s2[c & 3]); }
            const float sm = (s1[0] + s1[1]) + (s1[2] + s1[3]), sq = (s2[0] + s2[1]) + (s2[2] + s2[3]);
            const float mu = sm * (1.f / ET_C);
            const float var = fmaxf(fmaf(-mu, mu, sq * (1.f / ET_C)), 0.f);
            const bool valid = px < p.N;
            const float rstd = valid ? rsqrtf(var + ET_LN_EPS) : 0.f;
            const float nb = valid ? -mu * rstd : 0.f;
            // D2_READY also means the second GEMM has read the H operand of this buffer: t may overwrite it
            unsigned char* trow = tbuf + (size_t)b * ET_TILE_BYTES + row * 128;
#pragma unroll
            for (int oc = 0; oc < 16; ++oc) {
                uint4 pk;
                pk.x = pack_h2(fmaf(y[8 * oc], rstd, nb), fmaf(y[8 * oc + 1], rstd, nb));
                pk.y = pack_h2(fmaf(y[8 * oc + 2], rstd, nb), fmaf(y[8 * oc + 3], rstd, nb));
                pk.z = pack_h2(fmaf(y[8 * oc + 4], rstd, nb), fmaf(y[8 * oc + 5], rstd, nb));
                pk.w = pack_h2(fmaf(y[8 * oc + 6], rstd, nb), fmaf(y[8 * oc + 7], rstd, nb));
                *reinterpret_cast<uint4*>(trow + (oc >> 3) * ET_PANEL + xo[oc & 7]) = pk;
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
                unsigned char* dst = reinterpret_cast<unsigned char*>(p.tiles) +
                                     ((size_t)f * p.tiles_frame + tif) * ET_TILE_BYTES + sub * 4096;
                const unsigned char* src = tbuf + (size_t)b * ET_TILE_BYTES + sub * 4096;
                bulk_s2g(dst, src, 4096, pol);
                bulk_s2g(dst + ET_PANEL, src + ET_PANEL, 4096, pol);
                bulk_commit();
                // the store of the PREVIOUS tile (other buffer) has read its source: that buffer may be refilled
                if (i > 0) { bulk_wait_read<1>(); mbar_arrive(&bars[EB_T_FREE + (b ^ 1)]); }
            }
        }
        if (lane == 0) bulk_wait_read<0>();
    } else {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
        if (warp == 12) {
            if (lane == 0) {
                // ------------------------------- producer -------------------------------
                const uint64_t pol = l2_policy_evict_first(), pol_w = l2_policy_evict_last();
                mbar_arrive_expect_tx(&bars[EB_W_READY], ET_W1_BYTES + ET_W2_BYTES);
                bulk_g2s(w1s, p.w1, ET_W1_BYTES, &bars[EB_W_READY], pol_w);
                bulk_g2s(w2s, p.w2, ET_W2_BYTES, &bars[EB_W_READY], pol_w);
                for (int i = 0; i < my_tiles; ++i) {
                    const int g = (int)blockIdx.x + i * (int)gridDim.x;
                    const int f = g / p.tiles_frame, tif = g % p.tiles_frame;
                    const int b = i & 1;
                    mbar_wait(&bars[EB_IN_EMPTY + b], ((i >> 1) & 1) ^ 1);
                    if (tif * ET_TILE_PX < p.N) {
                        mbar_arrive_expect_tx(&bars[EB_IN_FULL + b], ET_IN_BYTES);
                        if (NHWC) {
                            tma_load_3d(inbuf + (size_t)b * ET_IN_BYTES, &tmap, 0, tif * ET_TILE_PX, f, &bars[EB_IN_FULL + b], pol);
                            tma_load_3d(inbuf + (size_t)b * ET_IN_BYTES + ET_IN_BYTES / 2, &tmap, 32, tif * ET_TILE_PX, f,
                                        &bars[EB_IN_FULL + b], pol);
                        } else {
                            tma_load_3d(inbuf + (size_t)b * ET_IN_BYTES, &tmap, tif * ET_TILE_PX, 0, f, &bars[EB_IN_FULL + b], pol);
                        }
                    } else {
                        mbar_arrive(&bars[EB_IN_FULL + b]);
                    }
                }
            }
        } else if (warp == 13) {
            // ------------------------------- MMA issuer -------------------------------
            const uint32_t idesc = umma_idesc_f16(128, 128);
            const uint32_t a1_u32 = smem_u32(a1buf), t_u32 = smem_u32(tbuf);
            const uint64_t dw1 = umma_smem_desc(smem_u32(w1s)), dw2 = umma_smem_desc(smem_u32(w2s));
            mbar_wait(&bars[EB_W_READY], 0);
            auto gemm2 = [&](int u) {
                const int ub = u & 1;
                mbar_wait(&bars[EB_A2_READY + ub], (u >> 1) & 1);
                mbar_wait(&bars[EB_D2_FREE + ub], ((u >> 1) & 1) ^ 1);
                tcgen05_fence_after();
                const uint64_t da = umma_smem_desc(t_u32 + (uint32_t)ub * ET_TILE_BYTES);
                if (elect_one()) {
#pragma unroll
                    for (int kb = 0; kb < 2; ++kb)
#pragma unroll
                        for (int k4 = 0; k4 < 4; ++k4)
                            umma_f16(tmem + 256u + 128u * (uint32_t)ub, da + (uint64_t)((kb * ET_PANEL + k4 * 32) >> 4),
                                     dw2 + (uint64_t)((kb * ET_PANEL + k4 * 32) >> 4), idesc, (kb | k4) != 0);
                    umma_commit(&bars[EB_D2_READY + ub]);
                }
                __syncwarp();
            };
            for (int i = 0; i < my_tiles; ++i) {
                const int b = i & 1;
                mbar_wait(&bars[EB_A1_READY + b], (i >> 1) & 1);
                tcgen05_fence_after();
                const uint64_t da = umma_smem_desc(a1_u32 + (uint32_t)b * ET_A1_BYTES);
                if (elect_one()) {
#pragma unroll
                    for (int k4 = 0; k4 < 4; ++k4)
                        umma_f16(tmem + 128u * (uint32_t)b, da + (uint64_t)((k4 * 32) >> 4), dw1 + (uint64_t)((k4 * 32) >> 4),
                                 idesc, k4 != 0);
                    umma_commit(&bars[EB_D1_READY + b]);
                    umma_commit(&bars[EB_A1_FREE + b]);
                }
                __syncwarp();
                if (i > 0) gemm2(i - 1);
            }
            if (my_tiles > 0) gemm2(my_tiles - 1);
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 13) tmem_dealloc(tmem, 512);
}

// ---------------------------------------------------------------------------------------------------------
// weight preparation: fold LN_64's affine into the first layer, pack both matrices as K-major swizzled fp16
// operand images, collapse the positional dense layer (grid = (y, x, 1-y, 1-x)) to 3 constants per channel
// ---------------------------------------------------------------------------------------------------------
__global__ void enc_tail_prep_kernel(const float* __restrict__ pos_w, const float* __restrict__ pos_b,
                                     const float* __restrict__ ln_w, const float* __restrict__ ln_b,
                                     const float* __restrict__ w1, const float* __restrict__ b1,
                                     const float* __restrict__ w2, const float* __restrict__ b2,
                                     __half* __restrict__ w1p, __half* __restrict__ w2p, float* __restrict__ par) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < ET_C * ET_CIN) {                         // W1'[n][k] = W1[n][k] * gamma[k]
        const int n = idx / ET_CIN, k = idx % ET_CIN;
        w1p[n * 64 + ((((k >> 3) ^ (n & 7)) << 3) | (k & 7))] = __float2half_rn(w1[idx] * ln_w[k]);
    } else if (idx < ET_C * ET_CIN + ET_C * ET_C) {
        const int j = idx - ET_C * ET_CIN;
        const int n = j / ET_C, k = j % ET_C;
        w2p[(k >> 6) * (ET_PANEL / 2) + n * 64 + (((((k & 63) >> 3) ^ (n & 7)) << 3) | (k & 7))] = __float2half_rn(w2[j]);
    } else if (idx < ET_C * ET_CIN + ET_C * ET_C + ET_C) {
        const int n = idx - (ET_C * ET_CIN + ET_C * ET_C);
        float a = b1[n];                                  // b1' = b1 + W1 beta
        for (int k = 0; k < ET_CIN; ++k) a = fmaf(w1[n * ET_CIN + k], ln_b[k], a);
        par[n] = a;
        par[128 + n] = b2[n];
    } else if (idx < ET_C * ET_CIN + ET_C * ET_C + ET_C + ET_CIN) {
        const int c = idx - (ET_C * ET_CIN + ET_C * ET_C + ET_C);
        const float w0 = pos_w[4 * c], w1_ = pos_w[4 * c + 1], w2_ = pos_w[4 * c + 2], w3 = pos_w[4 * c + 3];
        par[256 + 4 * c] = w0 - w2_;                      // y coefficient
        par[256 + 4 * c + 1] = w1_ - w3;                  // x coefficient
        par[256 + 4 * c + 2] = pos_b[c] + w2_ + w3;
        par[256 + 4 * c + 3] = 0.f;
    }
}

size_t enc_tail_workspace_bytes() { return (size_t)ET_W1_BYTES + ET_W2_BYTES + ET_PAR_FLOATS * 4; }

cudaError_t enc_tail_prep_launch(const float* pos_w, const float* pos_b, const float* ln_w, const float* ln_b,
                                 const float* w1, const float* b1, const float* w2, const float* b2, char* ws,
                                 cudaStream_t st) {
    const int total = ET_C * ET_CIN + ET_C * ET_C + ET_C + ET_CIN;
    enc_tail_prep_kernel<<<(total + 255) / 256, 256, 0, st>>>(
        pos_w, pos_b, ln_w, ln_b, w1, b1, w2, b2, reinterpret_cast<__half*>(ws),
        reinterpret_cast<__half*>(ws + ET_W1_BYTES), reinterpret_cast<float*>(ws + ET_W1_BYTES + ET_W2_BYTES));
    return cudaGetLastError();
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

cudaError_t enc_tail_launch(const float* cnn, long long frame_stride, int frames, int H, int W, void* tiles,
                            const char* ws, int sms, int tiles_frame, bool nhwc, cudaStream_t st) {
    static EncodeTiledFn enc = nullptr;
    if (enc == nullptr) {
        void* sym = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) != cudaSuccess ||
            qres != cudaDriverEntryPointSuccess)
            return cudaErrorNotSupported;
        enc = reinterpret_cast<EncodeTiledFn>(sym);
    }
    const int N = H * W;
    EncTailParams p{};
    p.tiles = reinterpret_cast<__half*>(tiles);
    p.w1 = reinterpret_cast<const __half*>(ws);
    p.w2 = reinterpret_cast<const __half*>(ws + ET_W1_BYTES);
    p.par = reinterpret_cast<const float*>(ws + ET_W1_BYTES + ET_W2_BYTES);
    p.frames = frames; p.N = N; p.W = W;
    p.tiles_frame = tiles_frame;                   // tiles per frame of the buffer (>= ceil(N / 128): every one is written)
    p.inv_hm1 = H > 1 ? 1.f / (float)(H - 1) : 0.f;
    p.inv_wm1 = W > 1 ? 1.f / (float)(W - 1) : 0.f;
    // [frames][64 channels][N pixels] fp32 (NCHW): box = 128 pixels x 64 channels, no swizzle (thread = pixel reads
    // consecutive addresses across the warp); pixels beyond N are zero-filled
    CUtensorMap tmap;
    const cuuint32_t estr[3] = {1u, 1u, 1u};
    CUresult cr;
    if (nhwc) {
        // [frames][N pixels][64 channels] fp32 (channels-last): box = 32 channels (128 B) x 128 pixels, 128-byte swizzle
        const cuuint64_t gdim[3] = {(cuuint64_t)ET_CIN, (cuuint64_t)N, (cuuint64_t)frames};
        const cuuint64_t gstride[2] = {(cuuint64_t)ET_CIN * 4, (cuuint64_t)frame_stride * 4};
        const cuuint32_t box[3] = {32u, (cuuint32_t)ET_TILE_PX, 1u};
        cr = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(cnn), gdim, gstride, box, estr,
                 CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                 CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    } else {
        const cuuint64_t gdim[3] = {(cuuint64_t)N, (cuuint64_t)ET_CIN, (cuuint64_t)frames};
        const cuuint64_t gstride[2] = {(cuuint64_t)N * 4, (cuuint64_t)frame_stride * 4};
        const cuuint32_t box[3] = {(cuuint32_t)ET_TILE_PX, (cuuint32_t)ET_CIN, 1u};
        cr = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(cnn), gdim, gstride, box, estr,
                 CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                 CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    }
    if (cr != CUDA_SUCCESS) return cudaErrorInvalidValue;
    auto kern = nhwc ? enc_tail_kernel<true> : enc_tail_kernel<false>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, ET_SMEM);
    if (e != cudaSuccess) return e;
    const int total = frames * p.tiles_frame;
    const int grid = total < sms ? total : sms;
    kern<<<grid, 512, ET_SMEM, st>>>(p, tmap);
    return cudaGetLastError();
}

}  // namespace sfb
