// Slot Attention streaming passes on the 5th-generation tensor cores (tcgen05 + TMEM), C = D = 128.
//
// Same operator and same workspace contract as sa_pass_kernel (sa_pass.cu; reference
// base_slots/models/savi.py:76-89, steve.py:43-55): per (frame, pixel-chunk) item it produces the partial
// sums  sum_n a[n,m] t[n,:],  sum_n a[n,m]  and  sum_n t[n,:]  that sa_update_kernel turns into the slot update.
// What changes is who does the work.  sa_pass_kernel runs one warp per 16-pixel tile through a serial chain
// (TMA wait -> LayerNorm with quad shuffles -> ldmatrix + mma.sync logits -> shuffle softmax -> movmatrix ->
// ldmatrix.trans + mma.sync aggregation) at 240 registers and 2 warps per scheduler: ncu shows it bound by
// dependent-instruction latency (35 % of the issue slots), not by HBM.  Here a 128-pixel tile is the unit:
//
//   producer warp   tensor-map TMA (cp.async.bulk.tensor, SWIZZLE_128B boxes of 32 pixels x 128 B) of the raw
//                   feature rows into a ring of 32-pixel stages; one 4 KB bulk copy of the frame's q~ operand
//                   per item.  Ragged tails are zero-filled by the TMA unit.
//   8 LN warps      ONE THREAD PER PIXEL: the whole 128-channel row goes to registers (conflict-free thanks to
//                   the TMA swizzle), mean / variance are plain in-thread sums (no shuffles),
//                   t = (x - mu) * rstd is written as the fp16 [128 px x 128 ch] operand tile in the canonical
//                   128-byte-swizzled layout (two 64-channel panels) and sent to the x^ ring with TMA bulk
//                   stores.  (LayerNorm's affine is folded into q~ and the slot update)
//   MMA warp        one elected thread:  logits[128 px x 16] (TMEM) = T (A, K-major) x [q~_hi ; q~_lo]^T (B), then
//                   U^T[128 ch x 16] (TMEM) += T^T (A, the SAME shared-memory tile read MN-major) x P (B),
//                   accumulated in TMEM over the 8 tiles of an item.
//   4 softmax warps ONE THREAD PER PIXEL (= TMEM lane): tcgen05.ld of the 16 logit columns, hi + lo + bias,
//                   softmax over the <= 8 slots entirely in registers, P = fp16(1024 a) as one 16-byte store
//                   (MN-major B operand, 8 slots per pixel); column 8 of P is the constant 1, so the
//                   aggregation also delivers sum_n t[n,:].  At item end they read U^T out of TMEM (thread =
//                   channel) and write the partials.
//
// Later iterations (sa_pass_tc_next_kernel) have no LayerNorm at all: the x^ ring holds ready-made operand
// tiles, one 32 KB bulk copy per tile feeds the tensor cores directly.
// Item order, the per-item partial layout and the fp16 rounding points (t, 1024 a) are those of sa_pass_kernel;
// sums are accumulated in a different order (TMEM), so results agree to fp32 rounding, not bit for bit.
#include "umma.cuh"
#include "sa_kernel.h"

#include <cuda.h>

namespace sfb {

namespace {

constexpr float TC_PSCALE = 1024.f;      // must match sa_update.cu (SA_PSCALE)
constexpr float TC_LN_EPS = 1e-5f;
constexpr int TC_C = 128;
constexpr int TC_TILE_PX = 128, TC_SUB_PX = 32;
constexpr int TC_PANEL_BYTES = TC_TILE_PX * 128;          // 128 pixel rows x 64 channels fp16
constexpr int TC_TILE_BYTES = 2 * TC_PANEL_BYTES;         // 32 KB
constexpr int TC_Q_BYTES = 4096;                          // [16 rows (hi 0-7, lo 8-15)] x 128 ch fp16, two panels
constexpr int TC_PPLANE = TC_TILE_PX * 16;                // one 8-column plane of P: 16 B per pixel

// TMEM columns (64 allocated): two logits buffers, two aggregation accumulators
constexpr uint32_t TC_COL_LOG = 0, TC_COL_ACC = 32;

// barrier slots (uint64 each)
enum {
    TB_FULL = 0,           // [8]  stage filled (TMA complete_tx)
    TB_EMPTY = 8,          // [8]  stage rows are in registers
    TB_TREADY = 16,        // [5]  operand tile written (LN warps / TMA)
    TB_TFREE = 21,         // [5]  aggregation MMAs have read the tile
    TB_LOG = 26,           // [2]  logits of the tile are in TMEM
    TB_PREADY = 28,        // [2]  P of the tile is in shared memory
    TB_ACC = 30,           // [2]  an item's accumulator is complete
    TB_ACCFREE = 32,       // [2]  ... and has been read out
    TB_QREADY = 34,        // [2]  q~ operand of an item landed
    TB_QFREE = 36,         // [2]  ... and its logits MMAs are done
    TB_PFREE = 38,         // [2]  the aggregation MMAs have read the P buffer
    TB_WORDS = 40
};

typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float lo, float hi) {
    f32x2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r;
}
__device__ __forceinline__ float lo2(f32x2 v) { float a, b; asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); return a; }
__device__ __forceinline__ float hi2(f32x2 v) { float a, b; asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); return b; }
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) { f32x2 r; asm("add.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
    f32x2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r;
}

// ---- shared-memory descriptors of the MN-major operands ---------------------------------------------------
// SWIZZLE_128B, MN-major (cute: ((8,n),(8,k)):((1,LBO),(8,SBO)) in 16-byte units): 64 MN elements contiguous,
// 8 K rows 128 B apart inside a swizzle atom, atoms LBO apart along MN and SBO apart along K.
__device__ __forceinline__ uint64_t umma_desc_mn_sw128(uint32_t addr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((addr & 0x3FFFFu) >> 4);
    d |= (uint64_t)(lbo >> 4) << 16;
    d |= (uint64_t)(sbo >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// no swizzle, MN-major (cute: ((1,n),(8,k)):((X,SBO),(1,LBO))): 8 MN elements (16 B) contiguous, 8 K rows 16 B
// apart inside a core matrix, core matrices LBO apart along K and SBO apart along MN.
__device__ __forceinline__ uint64_t umma_desc_mn_none(uint32_t addr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((addr & 0x3FFFFu) >> 4);
    d |= (uint64_t)(lbo >> 4) << 16;
    d |= (uint64_t)(sbo >> 4) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}
// instruction descriptor, kind::f16: fp16 operands, fp32 accumulate, optional MN-major A / B (bits 15 / 16)
__device__ __forceinline__ uint32_t umma_idesc_f16_major(int M, int N, bool a_mn, bool b_mn) {
    return (1u << 4) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(N >> 3) << 17) |
           ((uint32_t)(M >> 4) << 24);
}

// TMEM -> registers: this warp's 32 lanes x 16 consecutive columns
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 "
                 "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr));
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// tensor-map TMA: one box of the [frames][pixels][channels] feature tensor -> shared memory (SASS: UTMALDG)
__device__ __forceinline__ void tma_load_3d(void* dst_smem, const CUtensorMap* tmap, int c0, int c1, int c2,
                                            uint64_t* bar, uint64_t policy) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
                 " [%0], [%1, {%2, %3, %4}], [%5], %6;"
                 :: "r"(smem_u32(dst_smem)), "l"(tmap), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar)), "l"(policy)
                 : "memory");
}

// the same with a 4-D map [frames][128-byte channel chunks][pixels][channels of a chunk]: ONE instruction fetches all
// channel chunks of a 32-pixel stage (16 KB) -- the per-box cost of the TMA unit, not HBM, set the arrival rate of the raw
// stages with one 4 KB box per instruction
__device__ __forceinline__ void tma_load_4d(void* dst_smem, const CUtensorMap* tmap, int c0, int c1, int c2, int c3,
                                            uint64_t* bar, uint64_t policy) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
                 " [%0], [%1, {%2, %3, %4, %5}], [%6], %7;"
                 :: "r"(smem_u32(dst_smem)), "l"(tmap), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(bar)), "l"(policy)
                 : "memory");
}

// arrive whose issue depends on `dep` (a value derived from loaded data): the barrier is signalled only after
// the loads that produced `dep` have delivered their registers
__device__ __forceinline__ void mbar_arrive_after(uint64_t* bar, float dep) {
    uint32_t z;
    asm volatile("and.b32 %0, %1, 0;" : "=r"(z) : "r"(__float_as_uint(dep)));
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(bar) + z) : "memory");
}

// debug timeline (libsfb200_debug.so only): CTA 0, one lane per role stamps (role, tag, tile, clock) records
#ifdef SFB_DEBUG
struct TcProf {
    unsigned long long* buf;
    int n, cap;
    __device__ __forceinline__ TcProf(const SAPassParams& p, int role, bool on) {
        const bool act = on && p.prof != nullptr && blockIdx.x == 0 && p.prof_cap >= 8 * 512;
        buf = act ? p.prof + role * 512 : nullptr;
        n = 0; cap = 512;
    }
    __device__ __forceinline__ void mark(int tag, int tile) {
        if (buf != nullptr && n < cap)
            buf[n++] = ((unsigned long long)tag << 56) | ((unsigned long long)(tile & 0xffff) << 40) |
                       ((unsigned long long)clock64() & 0xFFFFFFFFFFull);
    }
};
#else
struct TcProf {
    __device__ __forceinline__ TcProf(const SAPassParams&, int, bool) {}
    __device__ __forceinline__ void mark(int, int) {}
};
#endif

// one lane polls, the warp sleeps at the warp barrier
__device__ __forceinline__ void mbar_wait_lane0(uint64_t* bar, uint32_t parity, int lane) {
    if (lane == 0) { while (!mbar_try_wait(bar, parity)) { } }
    __syncwarp();
}

template <bool FIRST, int EIN>
struct TcCfg {
    // first pass: two operand tile buffers (one per LayerNorm warp group) and one private raw stage per LayerNorm
    // warp (a warp's consecutive waits on ITS stage are consecutive phases of the barrier; a stage shared between
    // warps would let a warp wait on a parity that is two phases ahead of the barrier and pass at once);
    // later passes: five tile buffers fed by TMA directly
    static constexpr int NTB = FIRST ? 2 : 5;                             // operand tile buffers
    static constexpr int NST = FIRST ? 8 : 0;                             // raw 32-pixel stages
    static constexpr int STAGE_BYTES = TC_SUB_PX * TC_C * EIN;            // 16 KB fp32 / 8 KB bf16
    static constexpr int OFF_TILES = 0;
    static constexpr int OFF_Q = OFF_TILES + NTB * TC_TILE_BYTES;         // two q~ operands (1024-aligned)
    static constexpr int OFF_P = OFF_Q + 2 * TC_Q_BYTES;                  // [P0 | ones | P1 | ones]
    static constexpr int OFF_STAGES = OFF_P + 4 * TC_PPLANE;
    static constexpr int OFF_BARS = OFF_STAGES + NST * STAGE_BYTES;
    static constexpr int OFF_MISC = OFF_BARS + TB_WORDS * 8;              // tmem base, colsum scratch [2][4][8]
    static constexpr int SMEM = OFF_MISC + 16 + 2 * 4 * 8 * 4;
    static_assert(OFF_Q % 1024 == 0 && OFF_STAGES % 1024 == 0, "swizzled regions are 1024-byte aligned");
    static_assert(SMEM <= 232448, "tc pass: shared memory budget");
    static constexpr int SM_WARPS = 4;                                    // softmax / read-out warps (TMEM quadrants)
    static constexpr int LN_WARPS = FIRST ? 8 : 0;
    static constexpr int WARP_PROD = SM_WARPS + LN_WARPS;
    static constexpr int WARP_MMA = WARP_PROD + 1;
    static constexpr int THREADS = FIRST ? 512 : 256;                     // whole warpgroups (setmaxnreg)
};

// the items of this CTA: (frame, pixel chunk) pairs blockIdx.x, blockIdx.x + gridDim.x, ...
struct ItemIter {
    int items, my_items, tpi;      // tpi: 128-pixel tiles per item
    __device__ __forceinline__ ItemIter(const SAPassParams& p) {
        items = p.nframes * p.nchunk;
        my_items = (items > (int)blockIdx.x) ? (items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
        tpi = p.chunk_px / TC_TILE_PX;
    }
    __device__ __forceinline__ void locate(const SAPassParams& p, int il, int& f, int& chunk) const {
        const int idx = (int)blockIdx.x + il * (int)gridDim.x;
        const int item = p.reverse ? items - 1 - idx : idx;
        f = p.frame0 + item / p.nchunk;
        chunk = item % p.nchunk;
    }
};

// ---------------------------------------------------------------------------------------------------------
// MMA issuers (each warp runs its loop uniformly, one elected lane issues).  Two warps, two instruction streams:
// the logits of tile t need the operand tile, the aggregation of tile t needs its probabilities.  With one issuer
// and the order logits(t) -> aggregation(t-1) the release of a tile buffer (commit of the aggregation) queued
// behind the arrival of the NEXT tile, and the first pass ran at that cycle (4000 clocks per tile, measured with the
// role timeline); polling both barriers from one warp cost more than it gained (hot spin next to the softmax
// warps).  Each issuer now blocks on its own barrier.  What the single in-order stream used to guarantee is
// explicit: the logits issuer waits until the softmax warps have read the logits buffer it overwrites
// (PREADY of tile t-2), the softmax warps wait until the aggregation that read a P buffer is complete (PFREE).
// ---------------------------------------------------------------------------------------------------------
template <int NTB>
__device__ __forceinline__ void tc_logits_role(const SAPassParams& p, unsigned char* tiles, unsigned char* qbuf,
                                               uint64_t* bars, uint32_t tmem) {
    const ItemIter it(p);
    const uint32_t idesc_l = umma_idesc_f16_major(128, 16, false, false);
    const uint32_t tiles_u32 = smem_u32(tiles), q_u32 = smem_u32(qbuf);
    TcProf pf(p, 0, (threadIdx.x & 31) == 0);
    int tb = 0;                 // tile buffer of tile t
    uint32_t tbpar = 0;         // its fill parity
    int t = 0;
    for (int il = 0; il < it.my_items; ++il) {
        for (int j = 0; j < it.tpi; ++j, ++t) {
            if (j == 0) mbar_wait(&bars[TB_QREADY + (il & 1)], (il >> 1) & 1);
            if (t >= 2) mbar_wait(&bars[TB_PREADY + (t & 1)], ((t - 2) >> 1) & 1);    // logits buffer t & 1 has been read
            pf.mark(1, t);
            mbar_wait(&bars[TB_TREADY + tb], tbpar);
            pf.mark(2, t);
            tcgen05_fence_after();
            const uint64_t da = umma_smem_desc(tiles_u32 + (uint32_t)tb * TC_TILE_BYTES);
            const uint64_t db = umma_smem_desc(q_u32 + (uint32_t)(il & 1) * TC_Q_BYTES);
            const uint32_t dcol = tmem + TC_COL_LOG + 16u * (uint32_t)(t & 1);
            if (elect_one()) {
#pragma unroll
                for (int kb = 0; kb < 2; ++kb)
#pragma unroll
                    for (int k4 = 0; k4 < 4; ++k4)
                        umma_f16(dcol, da + (uint64_t)((kb * TC_PANEL_BYTES + k4 * 32) >> 4),
                                 db + (uint64_t)((kb * 2048 + k4 * 32) >> 4), idesc_l, (kb | k4) != 0);
                umma_commit(&bars[TB_LOG + (t & 1)]);
                if (j == it.tpi - 1) umma_commit(&bars[TB_QFREE + (il & 1)]);
            }
            __syncwarp();
            if (++tb == NTB) { tb = 0; tbpar ^= 1u; }
        }
    }
}

template <int NTB>
__device__ __forceinline__ void tc_agg_role(const SAPassParams& p, unsigned char* tiles, unsigned char* pbuf,
                                            uint64_t* bars, uint32_t tmem) {
    const ItemIter it(p);
    const uint32_t idesc_a = umma_idesc_f16_major(128, 16, true, true);
    const uint32_t tiles_u32 = smem_u32(tiles), p_u32 = smem_u32(pbuf);
    TcProf pf(p, 4, (threadIdx.x & 31) == 0);
    int ub = 0;                 // tile buffer of tile u
    int u = 0;
    for (int iu = 0; iu < it.my_items; ++iu) {
        for (int ju = 0; ju < it.tpi; ++ju, ++u) {
            if (ju == 0) mbar_wait(&bars[TB_ACCFREE + (iu & 1)], ((iu >> 1) & 1) ^ 1);
            pf.mark(3, u);
            mbar_wait(&bars[TB_PREADY + (u & 1)], (u >> 1) & 1);   // P of tile u is in shared memory (and its tile, hence, too)
            pf.mark(4, u);
            tcgen05_fence_after();
            const uint64_t da = umma_desc_mn_sw128(tiles_u32 + (uint32_t)ub * TC_TILE_BYTES, TC_PANEL_BYTES, 1024);
            const uint64_t db = umma_desc_mn_none(p_u32 + (uint32_t)(u & 1) * 2 * TC_PPLANE, 128, TC_PPLANE);
            const uint32_t dcol = tmem + TC_COL_ACC + 16u * (uint32_t)(iu & 1);
            if (elect_one()) {
#pragma unroll
                for (int k = 0; k < 8; ++k)         // 16 pixels per MMA: 2 KB of tile rows, 256 B of P
                    umma_f16(dcol, da + (uint64_t)((k * 2048) >> 4), db + (uint64_t)((k * 256) >> 4), idesc_a,
                             (ju | k) != 0);
                umma_commit(&bars[TB_TFREE + ub]);
                umma_commit(&bars[TB_PFREE + (u & 1)]);
                if (ju == it.tpi - 1) umma_commit(&bars[TB_ACC + (iu & 1)]);
            }
            __syncwarp();
            if (++ub == NTB) ub = 0;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// softmax + read-out warps (warp w owns TMEM lanes 32w .. 32w+31)
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tc_softmax_role(const SAPassParams& p, unsigned char* pbuf, float* csw,
                                                uint64_t* bars, uint32_t tmem, int warp, int lane, bool write_xsum) {
    const ItemIter it(p);
    const int K = p.K, N = p.N;
    const int row = 32 * warp + lane;                        // pixel inside the tile / channel at read-out
    const uint32_t lane_base = tmem + ((uint32_t)(32 * warp) << 16);
    TcProf pf(p, 1, threadIdx.x == 0);
    int t = 0;
    for (int il = 0; il < it.my_items; ++il) {
        int f, chunk;
        it.locate(p, il, f, chunk);
        float lb[8];
        {
            const float* bias = reinterpret_cast<const float*>(
                reinterpret_cast<const unsigned char*>(p.qt) + (size_t)f * (TC_Q_BYTES + 32) + TC_Q_BYTES);
#pragma unroll
            for (int s = 0; s < 8; ++s) lb[s] = __ldg(bias + s);
        }
        float cs[8];
#pragma unroll
        for (int s = 0; s < 8; ++s) cs[s] = 0.f;
        for (int j = 0; j < it.tpi; ++j, ++t) {
            const int b = t & 1;
            pf.mark(1, t);
            mbar_wait_lane0(&bars[TB_LOG + b], (t >> 1) & 1, lane);
            pf.mark(2, t);
            tcgen05_fence_after();
            float lg[16];
            tmem_ld16(lane_base + TC_COL_LOG + 16u * (uint32_t)b, lg);
            tmem_ld_wait();
            pf.mark(3, t);
            const int px = (chunk * it.tpi + j) * TC_TILE_PX + row;
            // one thread's softmax over the <= 8 slots is a dependent chain on a warp that has its scheduler to itself:
            // tree-shaped max / sum (depth 3 instead of 8), one ex2.approx per slot (slots >= K sit at -inf and come
            // out as exact zeros), the 1024 scale folded into the normaliser
            float e[8];
#pragma unroll
            for (int s = 0; s < 8; ++s) e[s] = (s < K) ? (lg[s] + lg[8 + s]) + lb[s] : -INFINITY;
            const float m = fmaxf(fmaxf(fmaxf(e[0], e[1]), fmaxf(e[2], e[3])), fmaxf(fmaxf(e[4], e[5]), fmaxf(e[6], e[7])));
#pragma unroll
            for (int s = 0; s < 8; ++s) asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e[s]) : "f"(e[s] - m));
            const float sum = ((e[0] + e[1]) + (e[2] + e[3])) + ((e[4] + e[5]) + (e[6] + e[7]));
            const float inv = (px < N) ? __fdividef(TC_PSCALE, sum) : 0.f;       // 1024 / sum
            if (p.seg_mask != nullptr && px < N) {
                float* mk = p.seg_mask + (size_t)f * K * N + px;
                const float inv1 = inv * (1.f / TC_PSCALE);
#pragma unroll
                for (int s = 0; s < 8; ++s)
                    if (s < K) mk[(size_t)s * N] = e[s] * inv1;
            }
            uint4 pk;
            {
                const __half2 h0 = __floats2half2_rn(e[0] * inv, e[1] * inv);
                const __half2 h1 = __floats2half2_rn(e[2] * inv, e[3] * inv);
                const __half2 h2 = __floats2half2_rn(e[4] * inv, e[5] * inv);
                const __half2 h3 = __floats2half2_rn(e[6] * inv, e[7] * inv);
                // column sums from the ROUNDED values (numerator and denominator of the update see the same a)
                const float2 f0 = __half22float2(h0), f1 = __half22float2(h1), f2 = __half22float2(h2), f3 = __half22float2(h3);
                cs[0] += f0.x; cs[1] += f0.y; cs[2] += f1.x; cs[3] += f1.y;
                cs[4] += f2.x; cs[5] += f2.y; cs[6] += f3.x; cs[7] += f3.y;
                pk.x = *reinterpret_cast<const uint32_t*>(&h0); pk.y = *reinterpret_cast<const uint32_t*>(&h1);
                pk.z = *reinterpret_cast<const uint32_t*>(&h2); pk.w = *reinterpret_cast<const uint32_t*>(&h3);
            }
            // the aggregation of tile t-2 (another issuer's instruction stream) has read this P buffer
            if (t >= 2) mbar_wait_lane0(&bars[TB_PFREE + b], ((t - 2) >> 1) & 1, lane);
            *reinterpret_cast<uint4*>(pbuf + (size_t)b * 2 * TC_PPLANE + row * 16) = pk;
            pf.mark(4, t);
            fence_proxy_async();            // P (generic proxy) -> tensor-core reads (async proxy)
            tcgen05_fence_before();         // the logits buffer has been read
            __syncwarp();
            if (lane == 0) mbar_arrive(&bars[TB_PREADY + b]);
            pf.mark(5, t);
        }
        // ---- item end: accumulator read-out (thread = channel) and column sums ----
        float* part = p.partials + ((size_t)f * p.nchunk + chunk) * p.pstride;
#pragma unroll
        for (int s = 0; s < 8; ++s) {
            float a = cs[s];
#pragma unroll
            for (int o = 16; o; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
            cs[s] = a;
        }
        float* my_csw = csw + (il & 1) * 32;
        if (lane == 0) {
#pragma unroll
            for (int s = 0; s < 8; ++s) my_csw[warp * 8 + s] = cs[s];
        }
        pf.mark(6, t);
        mbar_wait_lane0(&bars[TB_ACC + (il & 1)], (il >> 1) & 1, lane);
        pf.mark(7, t);
        tcgen05_fence_after();
        float u[16];
        tmem_ld16(lane_base + TC_COL_ACC + 16u * (uint32_t)(il & 1), u);
        tmem_ld_wait();
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars[TB_ACCFREE + (il & 1)]);
#pragma unroll
        for (int s = 0; s < 8; ++s) part[s * TC_C + row] = u[s];
        if (write_xsum) part[8 * TC_C + 8 + row] = u[8];          // sum_n t[n][row] (the constant-one column)
        named_bar_sync(2, 128);                                   // the four warps' column sums are in csw
        if (warp == 0 && lane < 8)
            part[8 * TC_C + lane] = (my_csw[lane] + my_csw[8 + lane]) + (my_csw[16 + lane] + my_csw[24 + lane]);
    }
}

// P planes 1 (columns 8..15) of both buffers: column 8 = 1, the rest 0
__device__ __forceinline__ void tc_init_ones(unsigned char* pbuf, int tid, int nthreads) {
    for (int i = tid; i < 2 * TC_TILE_PX; i += nthreads) {
        const int b = i / TC_TILE_PX, r = i % TC_TILE_PX;
        *reinterpret_cast<uint4*>(pbuf + (size_t)b * 2 * TC_PPLANE + TC_PPLANE + r * 16) = make_uint4(0x00003C00u, 0u, 0u, 0u);
    }
}

__device__ __forceinline__ void tc_init_bars(uint64_t* bars, int ntb, int nst, int ln_warps_per_tile) {
    for (int s = 0; s < nst; ++s) { mbar_init(&bars[TB_FULL + s], 1); mbar_init(&bars[TB_EMPTY + s], 1); }
    for (int s = 0; s < ntb; ++s) { mbar_init(&bars[TB_TREADY + s], ln_warps_per_tile); mbar_init(&bars[TB_TFREE + s], 1); }
    for (int s = 0; s < 2; ++s) {
        mbar_init(&bars[TB_LOG + s], 1);
        mbar_init(&bars[TB_PREADY + s], 4);
        mbar_init(&bars[TB_ACC + s], 1);
        mbar_init(&bars[TB_ACCFREE + s], 4);
        mbar_init(&bars[TB_QREADY + s], 1);
        mbar_init(&bars[TB_QFREE + s], 1);
        mbar_init(&bars[TB_PFREE + s], 1);
    }
    fence_mbar_init();
}

}  // namespace

// =========================================================================================================
// first pass: raw features -> LayerNorm -> operand tiles (+ x^ ring) -> tensor cores
// =========================================================================================================
template <int EIN>
__global__ void __launch_bounds__(512, 1) sa_pass_tc_first_kernel(const SAPassParams p,
                                                                 const __grid_constant__ CUtensorMap tmap) {
    using Cfg = TcCfg<true, EIN>;
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char* tiles = smem + Cfg::OFF_TILES;
    unsigned char* qbuf = smem + Cfg::OFF_Q;
    unsigned char* pbuf = smem + Cfg::OFF_P;
    unsigned char* stages = smem + Cfg::OFF_STAGES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::OFF_BARS);
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + Cfg::OFF_MISC);
    float* csw = reinterpret_cast<float*>(smem + Cfg::OFF_MISC + 16);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (tid == 0) tc_init_bars(bars, Cfg::NTB, Cfg::NST, 4);
    tc_init_ones(pbuf, tid, 512);
    if (warp == Cfg::WARP_MMA) tmem_alloc(tmem_ptr, 64);
    fence_proxy_async();
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem = *tmem_ptr;
    const ItemIter it(p);

    if (warp < Cfg::SM_WARPS) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 88;");
        tc_softmax_role(p, pbuf, csw, bars, tmem, warp, lane, true);
    } else if (warp < Cfg::WARP_PROD) {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 168;");
        // ------------------------------- LayerNorm warps: thread = pixel -------------------------------
        const int lw = warp - Cfg::SM_WARPS;
        const int total_sub = it.my_items * it.tpi * 4;
        const int N = p.N;
        uint32_t xo[8];                       // 16-byte chunk offsets of this thread's row under the 128B swizzle
#pragma unroll
        for (int j = 0; j < 8; ++j) xo[j] = (uint32_t)((j ^ (lane & 7)) << 4);
        const bool store_xhat = p.xhat != nullptr;
        // x^ stores: kept in L2 (evict_last) for the later pass to find -- unless the ring is far larger than L2 AND the
        // GPU is shared with the rollout (capped grid): then 400 MB of evict_last lines only push the rollout's weight
        // tiles out of L2
        const bool ring_fits_l2 = (size_t)p.xhat_frames * p.n16 * TC_C * 2 <= ((size_t)48 << 20);
        const uint64_t pol_x = (ring_fits_l2 || !p.cta_limited || p.xhat_keep) ? l2_policy_evict_last() : l2_policy_evict_first();
        TcProf pf(p, 2 + (lw >> 2), (lw & 3) == 0 && lane == 0);
        for (int n = lw; n < total_sub; n += 8) {
            static_assert(Cfg::NST == 8 && Cfg::NTB == 2, "one private stage per LayerNorm warp, one tile buffer per group");
            const int t = n >> 2, sub = n & 3, b = t & 1;
            const int s = lw;                                       // raw stage of sub-tile n: always this warp's own
            const uint32_t spar = (uint32_t)(n >> 3) & 1u;
            const int il = t / it.tpi, j = t - il * it.tpi;
            int f, chunk;
            it.locate(p, il, f, chunk);
            const int tile_in_frame = chunk * it.tpi + j;
            const int px = tile_in_frame * TC_TILE_PX + sub * TC_SUB_PX + lane;
            const unsigned char* stg = stages + (size_t)s * Cfg::STAGE_BYTES + lane * 128;
            pf.mark(1, t);
            mbar_wait_lane0(&bars[TB_FULL + s], spar, lane);
            pf.mark(2, t);
            f32x2 v[TC_C / 2];
            if (EIN == 4) {
#pragma unroll
                for (int q = 0; q < 4; ++q)
#pragma unroll
                    for (int jj = 0; jj < 8; ++jj) {
                        const ulonglong2 w = *reinterpret_cast<const ulonglong2*>(stg + q * 4096 + xo[jj]);
                        v[q * 16 + 2 * jj] = w.x; v[q * 16 + 2 * jj + 1] = w.y;
                    }
            } else {
#pragma unroll
                for (int q = 0; q < 2; ++q)
#pragma unroll
                    for (int jj = 0; jj < 8; ++jj) {
                        const uint4 w = *reinterpret_cast<const uint4*>(stg + q * 4096 + xo[jj]);
                        const uint32_t ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
                        for (int e = 0; e < 4; ++e)     // bf16 pair -> two fp32 (bit pattern shifted up)
                            v[q * 32 + 4 * jj + e] = pack2(__uint_as_float(ww[e] << 16), __uint_as_float(ww[e] & 0xffff0000u));
                    }
            }
            // pixels beyond N: a partially filled box is zero-filled by the TMA unit, but a sub-tile that starts beyond
            // N is never loaded -- its stage holds whatever shared memory held (possibly Inf / NaN patterns, and
            // 0 * NaN = NaN would reach the tensor cores): such rows are exact zeros
            if (px >= N) {
#pragma unroll
                for (int i = 0; i < TC_C / 2; ++i) v[i] = pack2(0.f, 0.f);
            }
            // statistics: plain in-thread sums over the row (4 independent chains each)
            f32x2 s2[4], q2[4];
#pragma unroll
            for (int a = 0; a < 4; ++a) { s2[a] = v[a]; q2[a] = fma2(v[a], v[a], pack2(0.f, 0.f)); }
#pragma unroll
            for (int i = 4; i < TC_C / 2; ++i) { s2[i & 3] = add2(s2[i & 3], v[i]); q2[i & 3] = fma2(v[i], v[i], q2[i & 3]); }
            const f32x2 st = add2(add2(s2[0], s2[1]), add2(s2[2], s2[3]));
            const f32x2 qt2 = add2(add2(q2[0], q2[1]), add2(q2[2], q2[3]));
            const float sm = lo2(st) + hi2(st), sq = lo2(qt2) + hi2(qt2);
            // the stage is free once every lane's row is in registers (sm depends on all of them)
            __syncwarp();
            if (lane == 0) mbar_arrive_after(&bars[TB_EMPTY + s], sm);
            const float mu = sm * (1.f / TC_C);
            const float var = fmaxf(fmaf(-mu, mu, sq * (1.f / TC_C)), 0.f);
            const bool valid = px < N;
            const float rstd = valid ? rsqrtf(var + TC_LN_EPS) : 0.f;     // rows beyond N become t = 0
            const float nb = valid ? -mu * rstd : 0.f;
            const f32x2 r2 = pack2(rstd, rstd), nb2 = pack2(nb, nb);
            // the tile buffer: free once the aggregation MMAs of tile t-2 have read it and this warp's own
            // x^ stores of tile t-2 (same rows) have read their source
            pf.mark(3, t);
            mbar_wait_lane0(&bars[TB_TFREE + b], ((t >> 1) & 1) ^ 1, lane);
            pf.mark(4, t);
            if (store_xhat) { if (lane == 0) bulk_wait_read<0>(); __syncwarp(); }
            pf.mark(5, t);
            unsigned char* trow = tiles + (size_t)b * TC_TILE_BYTES + (sub * TC_SUB_PX + lane) * 128;
#pragma unroll
            for (int oc = 0; oc < 16; ++oc) {         // 8 channels -> one 16-byte chunk of the operand row
                const f32x2 t0 = fma2(v[4 * oc], r2, nb2), t1 = fma2(v[4 * oc + 1], r2, nb2);
                const f32x2 t2 = fma2(v[4 * oc + 2], r2, nb2), t3 = fma2(v[4 * oc + 3], r2, nb2);
                uint4 pk;
                pk.x = pack_h2(lo2(t0), hi2(t0)); pk.y = pack_h2(lo2(t1), hi2(t1));
                pk.z = pack_h2(lo2(t2), hi2(t2)); pk.w = pack_h2(lo2(t3), hi2(t3));
                *reinterpret_cast<uint4*>(trow + (oc >> 3) * TC_PANEL_BYTES + xo[oc & 7]) = pk;
            }
            pf.mark(6, t);
            fence_proxy_async();              // operand rows (generic proxy) -> tensor cores / TMA store
            __syncwarp();
            pf.mark(7, t);
            if (lane == 0) {
                if (store_xhat) {
                    unsigned char* dst = reinterpret_cast<unsigned char*>(p.xhat) + (size_t)(f % p.xhat_frames) * p.xhat_fstride +
                                         (size_t)tile_in_frame * TC_TILE_BYTES + sub * 4096;
                    const unsigned char* src = tiles + (size_t)b * TC_TILE_BYTES + sub * 4096;
                    bulk_s2g(dst, src, 4096, pol_x);
                    bulk_s2g(dst + TC_PANEL_BYTES, src + TC_PANEL_BYTES, 4096, pol_x);
                    bulk_commit();
                }
                mbar_arrive(&bars[TB_TREADY + b]);
            }
            pf.mark(8, t);
        }
        if (store_xhat && lane == 0) bulk_wait_read<0>();
    } else {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 88;");
        if (warp == Cfg::WARP_PROD) {
            if (lane == 0) {
                // ------------------------------- TMA producer -------------------------------
                const uint64_t pol = l2_policy_evict_first();
                const uint64_t pol_q = l2_policy_evict_last();
                const int total_sub = it.my_items * it.tpi * 4;
                int s = 0;
                uint32_t spar = 0;
                for (int n = 0; n < total_sub; ++n) {
                    const int t = n >> 2, sub = n & 3;
                    const int il = t / it.tpi, j = t - il * it.tpi;
                    int f, chunk;
                    it.locate(p, il, f, chunk);
                    if (j == 0 && sub == 0) {
                        mbar_wait(&bars[TB_QFREE + (il & 1)], ((il >> 1) & 1) ^ 1);
                        mbar_arrive_expect_tx(&bars[TB_QREADY + (il & 1)], TC_Q_BYTES);
                        bulk_g2s(qbuf + (size_t)(il & 1) * TC_Q_BYTES,
                                 reinterpret_cast<const unsigned char*>(p.qt) + (size_t)f * (TC_Q_BYTES + 32), TC_Q_BYTES,
                                 &bars[TB_QREADY + (il & 1)], pol_q);
                    }
                    const int px0 = (chunk * it.tpi + j) * TC_TILE_PX + sub * TC_SUB_PX;
                    mbar_wait(&bars[TB_EMPTY + s], spar ^ 1u);
                    unsigned char* dst = stages + (size_t)s * Cfg::STAGE_BYTES;
                    if (px0 < p.N) {
                        mbar_arrive_expect_tx(&bars[TB_FULL + s], Cfg::STAGE_BYTES);
                        tma_load_4d(dst, &tmap, 0, px0, 0, f, &bars[TB_FULL + s], pol);
                    } else {
                        mbar_arrive(&bars[TB_FULL + s]);        // nothing to load: the LN warp writes zero rows
                    }
                    if (++s == Cfg::NST) { s = 0; spar ^= 1u; }
                }
            }
        } else if (warp == Cfg::WARP_MMA) {
            tc_logits_role<Cfg::NTB>(p, tiles, qbuf, bars, tmem);
        } else if (warp == Cfg::WARP_MMA + 1) {
            tc_agg_role<Cfg::NTB>(p, tiles, pbuf, bars, tmem);
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == Cfg::WARP_MMA) tmem_dealloc(tmem, 64);
}

// =========================================================================================================
// later passes: operand tiles straight from the x^ ring
// =========================================================================================================
__global__ void __launch_bounds__(256, 1) sa_pass_tc_next_kernel(const SAPassParams p) {
    using Cfg = TcCfg<false, 2>;
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char* tiles = smem + Cfg::OFF_TILES;
    unsigned char* qbuf = smem + Cfg::OFF_Q;
    unsigned char* pbuf = smem + Cfg::OFF_P;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::OFF_BARS);
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + Cfg::OFF_MISC);
    float* csw = reinterpret_cast<float*>(smem + Cfg::OFF_MISC + 16);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (tid == 0) tc_init_bars(bars, Cfg::NTB, 0, 1);
    tc_init_ones(pbuf, tid, 256);
    if (warp == Cfg::WARP_MMA) tmem_alloc(tmem_ptr, 64);
    fence_proxy_async();
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem = *tmem_ptr;
    const ItemIter it(p);

    if (warp < Cfg::SM_WARPS) {
        tc_softmax_role(p, pbuf, csw, bars, tmem, warp, lane, p.write_xsum != 0);
    } else if (warp == Cfg::WARP_PROD) {
        if (lane == 0) {
            const bool ring_fits_l2 = (size_t)p.xhat_frames * p.n16 * TC_C * 2 <= ((size_t)48 << 20);
            const uint64_t pol = ring_fits_l2 ? l2_policy_evict_last() : l2_policy_evict_first();
            const uint64_t pol_q = l2_policy_evict_last();
            int tb = 0;
            uint32_t tbpar = 0;
            for (int il = 0; il < it.my_items; ++il) {
                int f, chunk;
                it.locate(p, il, f, chunk);
                mbar_wait(&bars[TB_QFREE + (il & 1)], ((il >> 1) & 1) ^ 1);
                mbar_arrive_expect_tx(&bars[TB_QREADY + (il & 1)], TC_Q_BYTES);
                bulk_g2s(qbuf + (size_t)(il & 1) * TC_Q_BYTES,
                         reinterpret_cast<const unsigned char*>(p.qt) + (size_t)f * (TC_Q_BYTES + 32), TC_Q_BYTES,
                         &bars[TB_QREADY + (il & 1)], pol_q);
                for (int j = 0; j < it.tpi; ++j) {
                    const int tile_in_frame = chunk * it.tpi + j;
                    mbar_wait(&bars[TB_TFREE + tb], tbpar ^ 1u);
                    mbar_arrive_expect_tx(&bars[TB_TREADY + tb], TC_TILE_BYTES);
                    bulk_g2s(tiles + (size_t)tb * TC_TILE_BYTES,
                             reinterpret_cast<const unsigned char*>(p.xhat) + (size_t)(f % p.xhat_frames) * p.xhat_fstride +
                                 (size_t)tile_in_frame * TC_TILE_BYTES,
                             TC_TILE_BYTES, &bars[TB_TREADY + tb], pol);
                    if (++tb == Cfg::NTB) { tb = 0; tbpar ^= 1u; }
                }
            }
        }
    } else if (warp == Cfg::WARP_MMA) {
        tc_logits_role<Cfg::NTB>(p, tiles, qbuf, bars, tmem);
    } else if (warp == Cfg::WARP_MMA + 1) {
        tc_agg_role<Cfg::NTB>(p, tiles, pbuf, bars, tmem);
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == Cfg::WARP_MMA) tmem_dealloc(tmem, 64);
}

// =========================================================================================================
// host side
// =========================================================================================================
bool sa_pass_tc_supported(const SAPassParams& p, int C) {
    return C == TC_C && p.K >= 1 && p.K <= 8 && (p.chunk_px % TC_TILE_PX) == 0 && (p.n16 % TC_TILE_PX) == 0;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        void* sym = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(sym);
        tried = true;
    }
    return fn;
}

cudaError_t sa_pass_tc_launch(const SAPassParams& p, bool first, int sms, cudaStream_t st) {
    const int items = p.nframes * p.nchunk;
    const int grid = items < sms ? items : sms;
    if (!first) {
        using Cfg = TcCfg<false, 2>;
        cudaError_t e = cudaFuncSetAttribute(sa_pass_tc_next_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM);
        if (e != cudaSuccess) return e;
        sa_pass_tc_next_kernel<<<grid, Cfg::THREADS, Cfg::SMEM, st>>>(p);
        return cudaGetLastError();
    }
    EncodeTiledFn enc = encode_tiled_fn();
    if (enc == nullptr) return cudaErrorNotSupported;
    // [frames][pixels][channels] view of the feature grids: box = 32 pixels x 128 bytes of channels, SWIZZLE_128B;
    // pixel rows beyond N are zero-filled by the TMA unit
    const int ein = p.feat_esize;
    const int nfr = p.frame0 + p.nframes;
    CUtensorMap tmap;
    // dims (innermost first): channels of a 128-byte chunk, pixels, chunks, frames; the box is a whole 32-pixel stage
    // [chunks][32 px][128 B] -- the chunk dimension has the SMALLER stride (128 B) than the pixel dimension (C * ein)
    const cuuint64_t gdim[4] = {(cuuint64_t)(128 / ein), (cuuint64_t)p.N, (cuuint64_t)(TC_C * ein / 128), (cuuint64_t)nfr};
    const cuuint64_t gstride[3] = {(cuuint64_t)TC_C * ein, 128u, (cuuint64_t)p.feat_bstride * ein};
    const cuuint32_t box[4] = {(cuuint32_t)(128 / ein), (cuuint32_t)TC_SUB_PX, (cuuint32_t)(TC_C * ein / 128), 1u};
    const cuuint32_t estr[4] = {1u, 1u, 1u, 1u};
    const CUresult r = enc(&tmap, ein == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4,
                           const_cast<void*>(p.feats), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return cudaErrorInvalidValue;
    if (ein == 4) {
        using Cfg = TcCfg<true, 4>;
        auto kern = sa_pass_tc_first_kernel<4>;
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM);
        if (e != cudaSuccess) return e;
        kern<<<grid, Cfg::THREADS, Cfg::SMEM, st>>>(p, tmap);
    } else {
        using Cfg = TcCfg<true, 2>;
        auto kern = sa_pass_tc_first_kernel<2>;
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM);
        if (e != cudaSuccess) return e;
        kern<<<grid, Cfg::THREADS, Cfg::SMEM, st>>>(p, tmap);
    }
    return cudaGetLastError();
}

}  // namespace sfb
