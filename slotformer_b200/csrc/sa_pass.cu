// Slot Attention forward for sm_100a: batched streaming passes + batched slot updates.
//
// Replaces reference SlotAttention.forward (base_slots/models/savi.py:56-102) and
// SlotAttentionWMask.forward (base_slots/models/steve.py:19-73).
//
// Design (v0 -- one cluster per frame -- was latency bound by ~11 cluster
// syncs per frame, see profiles/r1_v0_timeline.txt):
//   * The K/V projections are never materialised.  With x^ = LN(x):
//        logits[n,m] = scale * <x^[n] Wk^T, q[m]>  = <x^[n], q~[m]>,   q~ = LNq(S) W_qk^T
//        updates[m]  = sum_n w[n,m] (x^[n] Wv^T)    = (sum_n w[n,m] x^[n]) Wv^T
//     so an iteration needs only x^ (N x C) and two K x C matrices; W_qk = scale*log2e*(Wq^T Wk)
//     and W_iv = W_ih Wv are folded once per call (sa_prep_kernel).
//   * sa_pass_kernel<FIRST>: persistent CTAs stream (frame, pixel-chunk) items.  Every warp runs
//     its own TMA pipeline (1-D bulk copies of 16-pixel tiles into a private mbarrier ring):
//     LayerNorm in place -> fp16 x^ tile (swizzled for ldmatrix) -> logits on tensor cores
//     (mma.sync, q~ split hi+lo) -> softmax over the slots with quad shuffles -> movmatrix
//     transpose -> U^T[C x 8] += x^T P on tensor cores.  The x^ tile is also written (as the
//     ready-made smem image) to a ring in global memory, so later iterations
//     (sa_pass_kernel<NEXT>) re-read 2 B/element (L2 resident when the ring is small) instead of
//     4 B/element from HBM.  +eps is applied analytically:
//        sum_n (a+eps) x^ = sum_n a x^ + eps sum_n x^ ,   sum_n (a+eps) = sum_n a + N eps.
//   * sa_update_kernel: GRU + residual MLP + next q~ for 16 rows (2 frames) per warp, batched
//     over all frames, tensor cores with 3-term fp16 splitting (hi*hi + lo*hi + hi*lo ~ fp32).
#include "common.cuh"
#include "sa_kernel.h"

namespace sfb {

// warps per CTA.  First pass: 8 (12 warps with 2-stage rings were measured SLOWER, 481 vs 433 us per Slot
// Attention call: with 8 KB fp32 stages ring depth matters more than warp count, and the kernel needs ~240
// registers).  Later passes stream 4 KB fp16 tiles and are bound by instruction latency (ncu: 35 % issue
// slots used, "wait" + short-scoreboard stalls), so they run 12 warps (3 per scheduler) with 4-stage rings.
constexpr int pass_warps(int C, bool first) { return (C == 128 && !first) ? 12 : 8; }
static constexpr float SA_PSCALE = 1024.f;   // probabilities are stored as fp16(1024 * a)
static constexpr float LN_EPS = 1e-5f;

bool sa_shape_supported(int C, int D, int DM) {
    return (C == 128 && D == 128 && DM == 256) || (C == 192 && D == 192 && DM == 384);
}

// ============================================================================================
// streaming pass
// ============================================================================================
// packed fp32x2 arithmetic (sm_100: FADD2 / FMUL2 / FFMA2) on 64-bit register pairs
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float lo, float hi) {
    f32x2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r;
}
__device__ __forceinline__ float lo2(f32x2 v) { float a, b; asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); return a; }
__device__ __forceinline__ float hi2(f32x2 v) { float a, b; asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); return b; }
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) { f32x2 r; asm("add.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) { f32x2 r; asm("mul.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
    f32x2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r;
}

template <int C, bool FIRST, int EIN>
struct PassCfg {
    static constexpr int NW = pass_warps(C, FIRST);
    static constexpr int KS = C / 16;
    static constexpr int ROWB = C * 2;                       // x^ row bytes
    static constexpr int XT_BYTES = 16 * ROWB;               // x^ tile (16 px) bytes
    static constexpr int STAGE_BYTES = FIRST ? 16 * C * EIN : XT_BYTES;   // EIN = input element bytes (4 fp32, 2 bf16)
    static constexpr int NST = (FIRST && EIN == 4) ? (C == 128 ? 3 : 2) : (C == 128 ? (FIRST ? 6 : 4) : 4);
    static constexpr int NQB = (C == 128) ? 2 : 1;           // q~ buffers (double-buffered if room)
    static constexpr int NREG = KS * 4;
    static constexpr int OFF_STAGES = 0;
    static constexpr int OFF_RED = NW * NST * STAGE_BYTES;
    static constexpr int RED_BYTES = (NW / 2) * NREG * 32 * 4;
    static constexpr int OFF_QF = OFF_RED + RED_BYTES;
    static constexpr int QF_BYTES = 2 * 8 * C * 2 + 32;      // hi + lo + 8 fp32 logit biases
    static constexpr int OFF_LN = OFF_QF + NQB * QF_BYTES;
    static constexpr int OFF_CSW = OFF_LN;
    static constexpr int OFF_BARS = OFF_CSW + NW * 8 * 4;
    static constexpr int SMEM = OFF_BARS + NW * NST * 8;
    static_assert(SMEM <= 232448, "pass kernel: shared memory budget");
};

// XS (first pass, K == 8 only): sum_n t[n] is accumulated explicitly.  With K <= 7 the otherwise idle 8th
// slot column of the probability operand is set to 1, so the aggregation MMAs deliver sum_n t[n] for free.
template <int C, bool FIRST, int EIN, bool XS>
__global__ void __launch_bounds__(pass_warps(C, FIRST) * 32, 1) sa_pass_kernel(const SAPassParams p) {
    using Cfg = PassCfg<C, FIRST, EIN>;
    constexpr int PASS_WARPS = Cfg::NW, PASS_THREADS = PASS_WARPS * 32;
    constexpr int KS = Cfg::KS, ROWB = Cfg::ROWB, XT_BYTES = Cfg::XT_BYTES;
    constexpr int STAGE_BYTES = Cfg::STAGE_BYTES, NST = Cfg::NST, NREG = Cfg::NREG, NQB = Cfg::NQB;

    extern __shared__ __align__(1024) unsigned char smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    unsigned char* my_stages = smem + Cfg::OFF_STAGES + warp * NST * STAGE_BYTES;
    float* red = reinterpret_cast<float*>(smem + Cfg::OFF_RED);
    __half* qf = reinterpret_cast<__half*>(smem + Cfg::OFF_QF);
    float* colsum_w = reinterpret_cast<float*>(smem + Cfg::OFF_CSW);
    uint64_t* my_bars = reinterpret_cast<uint64_t*>(smem + Cfg::OFF_BARS) + warp * NST;

    const int N = p.N, K = p.K;
    const int items = p.nframes * p.nchunk;
    const int tiles_chunk0 = p.chunk_px >> 4;
    const int nbw = (tiles_chunk0 - warp + PASS_WARPS - 1) / PASS_WARPS;  // 16-px tiles of this warp per item (warp w: tiles w, w+NW, ...)
    const int tiles_chunk = p.chunk_px >> 4;
    const int tiles_frame = p.nchunk * tiles_chunk;
    const int my_items = (items > (int)blockIdx.x) ? (items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
    const uint32_t total_tiles = (uint32_t)my_items * nbw;

    if (lane == 0) {
        for (int s = 0; s < NST; ++s) mbar_init(&my_bars[s], 1);
        fence_mbar_init();
    }
    __syncthreads();
    if (my_items == 0) return;

    // later passes re-read the x^ ring: keep it in L2 only if it can fit there
    const bool ring_fits_l2 = (size_t)p.xhat_frames * p.n16 * C * 2 <= ((size_t)48 << 20);
    const uint64_t pol = (FIRST || !ring_fits_l2) ? l2_policy_evict_first() : l2_policy_evict_last();

    auto issue = [&](uint32_t n) {
        if (n < total_tiles && lane == 0) {
            if (FIRST) bulk_wait_read<0>();      // the x^ store that sourced this stage has read it
            const int il = n / nbw, j = n % nbw;
            const int item = (int)blockIdx.x + il * (int)gridDim.x;
            const int fl = item / p.nchunk, c = item % p.nchunk;
            const int f = p.frame0 + fl;
            const int tb = c * tiles_chunk + warp + PASS_WARPS * j;
            const int px0 = tb * 16;
            int nvalid = N - px0;
            nvalid = nvalid < 0 ? 0 : (nvalid > 16 ? 16 : nvalid);
            const int s = n % NST;
            if (nvalid > 0) {
                const void* src;
                uint32_t bytes;
                if (FIRST) {
                    src = reinterpret_cast<const unsigned char*>(p.feats) +
                          ((size_t)f * p.feat_bstride + (size_t)px0 * C) * EIN;
                    bytes = (uint32_t)nvalid * C * EIN;
                } else {
                    src = p.xhat + ((size_t)(f % p.xhat_frames) * tiles_frame + tb) * (16 * C);
                    bytes = XT_BYTES;
                }
                mbar_arrive_expect_tx(&my_bars[s], bytes);
                bulk_g2s(my_stages + s * STAGE_BYTES, src, bytes, &my_bars[s], pol);
            } else {
                mbar_arrive(&my_bars[s]);
            }
        }
    };
    auto load_qf = [&](int il, int buf, bool async) {
        const int item = (int)blockIdx.x + il * (int)gridDim.x;
        const int f = p.frame0 + item / p.nchunk;
        const uint4* src = reinterpret_cast<const uint4*>(p.qt + (size_t)f * (Cfg::QF_BYTES / 2));
        uint4* dst = reinterpret_cast<uint4*>(qf + (size_t)buf * (Cfg::QF_BYTES / 2));
        for (int i = tid; i < Cfg::QF_BYTES / 16; i += PASS_THREADS) {
            if (async) cp_async16(dst + i, src + i);
            else dst[i] = __ldg(src + i);
        }
        if (async) cp_async_commit();
    };

    const int g = lane >> 2, t4 = lane & 3;
    const int pxi = lane >> 3, ch8 = lane & 7;

#pragma unroll 1
    for (uint32_t n0 = 0; n0 + 1 < (uint32_t)NST; ++n0) issue(n0);
    load_qf(0, 0, false);
    __syncthreads();

    uint32_t n = 0;
#pragma unroll 1
    for (int il = 0; il < my_items; ++il) {
        const int item = (int)blockIdx.x + il * (int)gridDim.x;
        const int fl = item / p.nchunk, chunk = item % p.nchunk;
        const int f = p.frame0 + fl;
        const __half* qh = qf + (size_t)((NQB == 2) ? (il & 1) : 0) * (Cfg::QF_BYTES / 2);
        const __half* ql = qh + 8 * C;
        if (NQB == 2 && il + 1 < my_items) load_qf(il + 1, (il + 1) & 1, true);

        uint32_t bq_hi[KS][2], bq_lo[KS][2];
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) {
            const int c0 = 16 * ks + 2 * t4;
            bq_hi[ks][0] = *reinterpret_cast<const uint32_t*>(qh + g * C + c0);
            bq_hi[ks][1] = *reinterpret_cast<const uint32_t*>(qh + g * C + c0 + 8);
            bq_lo[ks][0] = *reinterpret_cast<const uint32_t*>(ql + g * C + c0);
            bq_lo[ks][1] = *reinterpret_cast<const uint32_t*>(ql + g * C + c0 + 8);
        }
        const float lb0 = reinterpret_cast<const float*>(ql + 8 * C)[2 * t4];
        const float lb1 = reinterpret_cast<const float*>(ql + 8 * C)[2 * t4 + 1];
        float uacc[KS][4];
#pragma unroll
        for (int cb = 0; cb < KS; ++cb) { uacc[cb][0] = uacc[cb][1] = uacc[cb][2] = uacc[cb][3] = 0.f; }
        float cs0 = 0.f, cs1 = 0.f;
        f32x2 xs2[C / 16];
#pragma unroll
        for (int i = 0; i < C / 16; ++i) xs2[i] = pack2(0.f, 0.f);
        const bool s0ok = (2 * t4) < K, s1ok = (2 * t4 + 1) < K;

#pragma unroll 1
        for (int j = 0; j < nbw; ++j, ++n) {
            issue(n + NST - 1);
            const int s = n % NST;
            unsigned char* stg = my_stages + s * STAGE_BYTES;
            const int tb = chunk * tiles_chunk + warp + PASS_WARPS * j;
            const int px0 = tb * 16;
            int nvalid = N - px0;
            nvalid = nvalid < 0 ? 0 : (nvalid > 16 ? 16 : nvalid);
            mbar_wait(&my_bars[s], (n / NST) & 1);
            if (nvalid > 0 && !(p.dbg & 4)) {
                if (FIRST && !(p.dbg & 8)) {
                    // ---- normalise 16 pixels in place: fp32 rows -> swizzled fp16 t = (x-mu)*rstd ----
                    // (the LayerNorm affine is folded into q~ and into the slot update; t rows 8h..8h+7
                    //  land on raw rows 4h..4h+3, which are already in registers)
                    if (nvalid < 16) {
                        // ragged tail: zero the missing raw rows so they normalise to t = 0
                        float4* z = reinterpret_cast<float4*>(stg + (size_t)nvalid * C * EIN);
                        for (int i = lane; i < (16 - nvalid) * (C * EIN / 16); i += 32) z[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                        __syncwarp();
                    }
                    {
                        // all 16 raw rows are loaded before anything is stored back (the fp16 t rows alias the
                        // raw rows), so the four row groups are independent and their latencies overlap
                        f32x2 v[4][C / 16];
#pragma unroll
                        for (int rr = 0; rr < 4; ++rr) {
                            const int row = 4 * rr + pxi;
                            const unsigned char* tp = stg + (size_t)(row * C + 4 * ch8) * EIN;
#pragma unroll
                            for (int i = 0; i < C / 32; ++i) {
                                if (EIN == 4) {
                                    const ulonglong2 q = *reinterpret_cast<const ulonglong2*>(tp + 32 * i * 4);
                                    v[rr][2 * i] = q.x; v[rr][2 * i + 1] = q.y;
                                } else {            // bf16: the fp32 bit pattern is the 16 bits shifted up
                                    const uint2 q = *reinterpret_cast<const uint2*>(tp + 32 * i * 2);
                                    v[rr][2 * i] = pack2(__uint_as_float(q.x << 16), __uint_as_float(q.x & 0xffff0000u));
                                    v[rr][2 * i + 1] = pack2(__uint_as_float(q.y << 16), __uint_as_float(q.y & 0xffff0000u));
                                }
                            }
                        }
                        __syncwarp();
                        // one sweep: sum and sum of squares together (one shuffle phase instead of two), then
                        // t = x * rstd - mu * rstd as a single fused multiply-add per pair
                        float sm[4], sq[4];
#pragma unroll
                        for (int rr = 0; rr < 4; ++rr) {
                            f32x2 s2 = v[rr][0], q2 = mul2(v[rr][0], v[rr][0]);
#pragma unroll
                            for (int i = 1; i < C / 16; ++i) { s2 = add2(s2, v[rr][i]); q2 = fma2(v[rr][i], v[rr][i], q2); }
                            sm[rr] = lo2(s2) + hi2(s2);
                            sq[rr] = lo2(q2) + hi2(q2);
                        }
#pragma unroll
                        for (int o = 1; o <= 4; o <<= 1) {
#pragma unroll
                            for (int rr = 0; rr < 4; ++rr) {
                                sm[rr] += __shfl_xor_sync(0xffffffffu, sm[rr], o);
                                sq[rr] += __shfl_xor_sync(0xffffffffu, sq[rr], o);
                            }
                        }
#pragma unroll
                        for (int rr = 0; rr < 4; ++rr) {
                            const int row = 4 * rr + pxi;
                            const float mu = sm[rr] * (1.f / C);
                            const float var = fmaxf(fmaf(-mu, mu, sq[rr] * (1.f / C)), 0.f);
                            const float rstd = rsqrtf(var + LN_EPS);
                            const f32x2 r2 = pack2(rstd, rstd);
                            const float nb = -mu * rstd;
                            const f32x2 nb2 = pack2(nb, nb);
                            unsigned char* rowp = stg + row * ROWB + (ch8 & 1) * 8;
#pragma unroll
                            for (int i = 0; i < C / 32; ++i) {
                                const f32x2 t0 = fma2(v[rr][2 * i], r2, nb2), t1 = fma2(v[rr][2 * i + 1], r2, nb2);
                                if (XS) { xs2[2 * i] = add2(xs2[2 * i], t0); xs2[2 * i + 1] = add2(xs2[2 * i + 1], t1); }
                                const int chk = ((ch8 >> 1) + 4 * i) ^ (row & 7);
                                uint2 pk; pk.x = pack_h2(lo2(t0), hi2(t0)); pk.y = pack_h2(lo2(t1), hi2(t1));
                                *reinterpret_cast<uint2*>(rowp + chk * 16) = pk;
                            }
                        }
                    }
                    __syncwarp();
                    if (p.xhat != nullptr && !(p.dbg & 2)) {
                        // ready-made smem image of the tile -> x^ ring in global memory: one TMA bulk store issued by
                        // lane 0 (no register round trip; the warp goes straight on to the tensor-core part)
                        fence_proxy_async();          // the generic-proxy writes of t precede the async-proxy read
                        __syncwarp();
                        if (lane == 0) {
                            bulk_s2g(p.xhat + ((size_t)(f % p.xhat_frames) * tiles_frame + tb) * (16 * C), stg, XT_BYTES,
                                     ring_fits_l2 ? l2_policy_evict_last() : l2_policy_evict_first());
                            bulk_commit();
                        }
                    }
                }
                if (p.dbg & 16) { __syncwarp(); continue; }      // (timing experiment: LayerNorm only)
                const uint32_t tile_u32 = smem_u32(stg);
                // ---- logits: 16 pixels x 8 slots, log2 domain (scale folded into q~) ----
                // four independent accumulator chains (hi/lo x even/odd k-step): legacy HMMA has a long
                // dependent-issue latency on sm_100, so short chains matter more than instruction count
                float lgA[4] = {lb0, lb1, lb0, lb1}, lgB[4] = {0.f, 0.f, 0.f, 0.f};
                float lgC[4] = {0.f, 0.f, 0.f, 0.f}, lgD[4] = {0.f, 0.f, 0.f, 0.f};
                {
                    const int row = (lane & 7) + ((lane >> 3) & 1) * 8;
                    const uint32_t rowa = tile_u32 + row * ROWB;
#pragma unroll
                    for (int ks = 0; ks < KS; ks += 2) {
                        uint32_t a0[4], a1[4];
                        ldsm_x4(a0, rowa + (((2 * ks + (lane >> 4)) ^ (row & 7)) << 4));
                        ldsm_x4(a1, rowa + (((2 * ks + 2 + (lane >> 4)) ^ (row & 7)) << 4));
                        mma_f16(lgA, a0, bq_hi[ks][0], bq_hi[ks][1]);
                        mma_f16(lgB, a0, bq_lo[ks][0], bq_lo[ks][1]);
                        mma_f16(lgC, a1, bq_hi[ks + 1][0], bq_hi[ks + 1][1]);
                        mma_f16(lgD, a1, bq_lo[ks + 1][0], bq_lo[ks + 1][1]);
                    }
                }
                float lg[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) lg[e] = (lgA[e] + lgC[e]) + (lgB[e] + lgD[e]);
                const int pxa = px0 + g, pxb = pxa + 8;
                float pa0, pa1, pb0, pb1;
                {
                    float ma = fmaxf(s0ok ? lg[0] : -INFINITY, s1ok ? lg[1] : -INFINITY);
                    float mb = fmaxf(s0ok ? lg[2] : -INFINITY, s1ok ? lg[3] : -INFINITY);
                    ma = fmaxf(ma, __shfl_xor_sync(0xffffffffu, ma, 1));
                    mb = fmaxf(mb, __shfl_xor_sync(0xffffffffu, mb, 1));
                    ma = fmaxf(ma, __shfl_xor_sync(0xffffffffu, ma, 2));
                    mb = fmaxf(mb, __shfl_xor_sync(0xffffffffu, mb, 2));
                    const float ea0 = s0ok ? exp2f(lg[0] - ma) : 0.f, ea1 = s1ok ? exp2f(lg[1] - ma) : 0.f;
                    const float eb0 = s0ok ? exp2f(lg[2] - mb) : 0.f, eb1 = s1ok ? exp2f(lg[3] - mb) : 0.f;
                    float sa = ea0 + ea1, sb = eb0 + eb1;
                    sa += __shfl_xor_sync(0xffffffffu, sa, 1);
                    sb += __shfl_xor_sync(0xffffffffu, sb, 1);
                    sa += __shfl_xor_sync(0xffffffffu, sa, 2);
                    sb += __shfl_xor_sync(0xffffffffu, sb, 2);
                    const float ia = (pxa < N) ? __fdividef(1.f, sa) : 0.f;
                    const float ib = (pxb < N) ? __fdividef(1.f, sb) : 0.f;
                    pa0 = ea0 * ia; pa1 = ea1 * ia; pb0 = eb0 * ib; pb1 = eb1 * ib;
                }
                if (p.seg_mask != nullptr) {
                    float* mk = p.seg_mask + (size_t)f * K * N;
                    if (pxa < N) {
                        if (s0ok) mk[(size_t)(2 * t4) * N + pxa] = pa0;
                        if (s1ok) mk[(size_t)(2 * t4 + 1) * N + pxa] = pa1;
                    }
                    if (pxb < N) {
                        if (s0ok) mk[(size_t)(2 * t4) * N + pxb] = pb0;
                        if (s1ok) mk[(size_t)(2 * t4 + 1) * N + pxb] = pb1;
                    }
                }
                if (FIRST && !XS && t4 == 3) { pa1 = 1.f / SA_PSCALE; pb1 = 1.f / SA_PSCALE; }   // slot column 7 := 1
                const __half2 ha = __floats2half2_rn(pa0 * SA_PSCALE, pa1 * SA_PSCALE);
                const __half2 hb = __floats2half2_rn(pb0 * SA_PSCALE, pb1 * SA_PSCALE);
                {
                    const float2 fa = __half22float2(ha), fb = __half22float2(hb);
                    cs0 += fa.x + fb.x; cs1 += fa.y + fb.y;   // column sums from the ROUNDED values
                }
                const uint32_t b0 = movmatrix_t(*reinterpret_cast<const uint32_t*>(&ha));
                const uint32_t b1 = movmatrix_t(*reinterpret_cast<const uint32_t*>(&hb));
                // ---- aggregation U^T[16 ch x 8 slots] += x^T[16 ch x 16 px] * P[16 px x 8 slots] ----
                {
                    const int row = (lane & 7) + (lane >> 4) * 8;
                    const uint32_t rowa = tile_u32 + row * ROWB;
#pragma unroll
                    for (int cb = 0; cb < KS; cb += 2) {
                        uint32_t a0[4], a1[4];
                        ldsm_x4_t(a0, rowa + (((2 * cb + ((lane >> 3) & 1)) ^ (row & 7)) << 4));
                        ldsm_x4_t(a1, rowa + (((2 * cb + 2 + ((lane >> 3) & 1)) ^ (row & 7)) << 4));
                        mma_f16(uacc[cb], a0, b0, b1);
                        mma_f16(uacc[cb + 1], a1, b0, b1);
                    }
                }
            }
            if (!(p.dbg & 1)) fence_proxy_async();   // generic-proxy accesses of this stage precede its next TMA fill
            __syncwarp();
        }

        // ================= item end: reduce the 8 warps' partials, write them out =================
        float* part = p.partials + ((size_t)f * p.nchunk + chunk) * p.pstride;
        cs0 += __shfl_xor_sync(0xffffffffu, cs0, 4);  cs1 += __shfl_xor_sync(0xffffffffu, cs1, 4);
        cs0 += __shfl_xor_sync(0xffffffffu, cs0, 8);  cs1 += __shfl_xor_sync(0xffffffffu, cs1, 8);
        cs0 += __shfl_xor_sync(0xffffffffu, cs0, 16); cs1 += __shfl_xor_sync(0xffffffffu, cs1, 16);
        if (lane < 4) { colsum_w[warp * 8 + 2 * lane] = cs0; colsum_w[warp * 8 + 2 * lane + 1] = cs1; }
        if (FIRST && XS) {
            float xs[C / 32][4];
#pragma unroll
            for (int i = 0; i < C / 32; ++i) {
                xs[i][0] = lo2(xs2[2 * i]); xs[i][1] = hi2(xs2[2 * i]);
                xs[i][2] = lo2(xs2[2 * i + 1]); xs[i][3] = hi2(xs2[2 * i + 1]);
            }
#pragma unroll
            for (int i = 0; i < C / 32; ++i)
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    float a = xs[i][e];
                    a += __shfl_xor_sync(0xffffffffu, a, 8);
                    a += __shfl_xor_sync(0xffffffffu, a, 16);
                    xs[i][e] = a;
                }
            if (lane < 8) {
#pragma unroll
                for (int i = 0; i < C / 32; ++i)
                    *reinterpret_cast<float4*>(red + warp * C + 4 * lane + 32 * i) =
                        make_float4(xs[i][0], xs[i][1], xs[i][2], xs[i][3]);
            }
            __syncthreads();
            if (tid < C) {
                float a = 0.f;
#pragma unroll
                for (int w8 = 0; w8 < PASS_WARPS; ++w8) a += red[w8 * C + tid];
                part[8 * C + 8 + tid] = a;
            }
            __syncthreads();
        }
        {
            auto put = [&](int slot) {
#pragma unroll
                for (int cb = 0; cb < KS; ++cb)
#pragma unroll
                    for (int e = 0; e < 4; ++e) red[(slot * NREG + cb * 4 + e) * 32 + lane] = uacc[cb][e];
            };
            auto add = [&](int slot) {
#pragma unroll
                for (int cb = 0; cb < KS; ++cb)
#pragma unroll
                    for (int e = 0; e < 4; ++e) uacc[cb][e] += red[(slot * NREG + cb * 4 + e) * 32 + lane];
            };
            // tree over PASS_WARPS warps (8: 4-2-1, 12: 6-3-(1+2)); a warp only overwrites a slot it
            // has already consumed itself
            constexpr int H1 = PASS_WARPS / 2, H2 = PASS_WARPS / 4;
            if (warp >= H1) put(warp - H1);
            __syncthreads();
            if (warp < H1) add(warp);
            if (warp >= H2 && warp < H1) put(warp);
            __syncthreads();
            if (PASS_WARPS == 8) {
                if (warp < 2) add(2 + warp);
                if (warp == 1) put(1);
            } else {
                if (warp < 3) add(3 + warp);
                if (warp == 1 || warp == 2) put(warp);
            }
            __syncthreads();
            if (warp == 0) {
                add(1);
                if (PASS_WARPS == 12) add(2);
#pragma unroll
                for (int cb = 0; cb < KS; ++cb)
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const int c = 16 * cb + g + 8 * (e >> 1);
                        const int slot = 2 * t4 + (e & 1);
                        part[slot * C + c] = uacc[cb][e];
                        if (FIRST && !XS && slot == 7) part[8 * C + 8 + c] = uacc[cb][e];   // sum_n t[n][c]
                    }
            } else if (warp == 1 && lane < 8) {
                float a = 0.f;
#pragma unroll
                for (int w8 = 0; w8 < PASS_WARPS; ++w8) a += colsum_w[w8 * 8 + lane];
                part[8 * C + lane] = a;
            }
        }
        if (NQB == 2) {
            cp_async_wait_all();
        } else if (il + 1 < my_items) {
            __syncthreads();                 // everyone finished reading the single q~ buffer
            load_qf(il + 1, 0, false);
        }
        __syncthreads();
    }
    if (FIRST && lane == 0) bulk_wait_read<0>();   // no x^ store may still be reading this CTA's shared memory
}

template <int C, bool FIRST, int EIN, bool XS>
static cudaError_t pass_launch_x(const SAPassParams& p, int sms, cudaStream_t st) {
    using Cfg = PassCfg<C, FIRST, EIN>;
    auto kern = sa_pass_kernel<C, FIRST, EIN, XS>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM);
    if (e != cudaSuccess) return e;
    const int items = p.nframes * p.nchunk;
    const int grid = items < sms ? items : sms;
    kern<<<grid, Cfg::NW * 32, Cfg::SMEM, st>>>(p);
    return cudaGetLastError();
}

template <int C, bool FIRST, int EIN>
static cudaError_t pass_launch_t(const SAPassParams& p, int sms, cudaStream_t st) {
    if (FIRST && p.K == 8) return pass_launch_x<C, FIRST, EIN, true>(p, sms, st);
    return pass_launch_x<C, FIRST, EIN, false>(p, sms, st);
}

cudaError_t sa_pass_launch(const SAPassParams& p, int C, bool first, int sms, int smem_limit, cudaStream_t st) {
    (void)smem_limit;
    {
        // warp-pair first pass (sa_pass_split.cu): 9 % faster when the grid is capped (SM-bound, 633 -> 576 us per
        // Slot Attention call at 84 CTAs), 1 % slower with every SM (HBM-bound); SFB_SA_SPLIT_ON / _OFF override
        const bool want = p.split >= 0 ? (p.split != 0) : (p.cta_limited != 0);
        if (first && want && sa_pass_split_supported(p, C)) return sa_pass_split_launch(p, sms, st);
    }
    const bool bf16 = p.feat_esize == 2;
    if (C == 128) {
        if (!first) return pass_launch_t<128, false, 4>(p, sms, st);
        return bf16 ? pass_launch_t<128, true, 2>(p, sms, st) : pass_launch_t<128, true, 4>(p, sms, st);
    }
    if (!first) return pass_launch_t<192, false, 4>(p, sms, st);
    return bf16 ? pass_launch_t<192, true, 2>(p, sms, st) : pass_launch_t<192, true, 4>(p, sms, st);
}

}  // namespace sfb
