// Slot Attention first pass with the two halves of a pixel tile on DIFFERENT warps (C = 128, K <= 7, no mask).
//
// sa_pass_kernel<FIRST> does, per 16-pixel tile and on one warp: TMA wait -> LayerNorm -> fp16 tile -> logits /
// softmax / aggregation on tensor cores.  Measured at the batch pipeline's 84-CTA cap that chain is additive
// (streaming 150 us + LayerNorm 158 us + tensor-core part 100 us of a 350 us pass): a warp's tile time is the sum of
// the three, and more warps of the same kind do not fit (240 registers, 8 KB fp32 stages).  Here every pixel-tile
// stream is served by a PAIR of warps:
//   * warp p (front, 8 of them): its own 2-stage fp32 TMA ring; per tile it pulls the 16 raw rows into registers,
//     re-arms the stage with the tile two ahead at once, normalises (single-sweep statistics), writes the fp16 t
//     tile into one of two 4 KB slots, sends it to the x^ ring with a TMA bulk store and signals the back warp;
//   * warp 8 + p (back): waits for the slot, runs logits (q~ split hi + lo) -> softmax over slots -> U^T += t^T P
//     exactly as sa_pass_kernel does, and hands the slot back.
// The fp32 stage is free as soon as its rows are in registers, so two loads per pair stay in flight while the
// front warp computes, and the LayerNorm of tile n+1 overlaps the tensor-core work of tile n.  Items, partial-sum
// layout, summation order inside a warp and the cross-warp tree are those of sa_pass_kernel (the back warps take
// the roles of its 8 warps), so results are bit-identical to it.
#include "common.cuh"
#include "sa_kernel.h"

namespace sfb {

namespace {

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

constexpr int SP_C = 128, SP_PAIRS = 8, SP_THREADS = 2 * SP_PAIRS * 32;
constexpr int SP_KS = SP_C / 16, SP_ROWB = SP_C * 2, SP_XT = 16 * SP_ROWB;     // fp16 tile: 4 KB
constexpr int SP_NREG = SP_KS * 4;
constexpr float SP_PSCALE = 1024.f, SP_LN_EPS = 1e-5f;

template <int EIN>
struct SplitCfg {
    static constexpr int STAGE = 16 * SP_C * EIN;                     // raw tile: 8 KB fp32 / 4 KB bf16
    static constexpr int NST = 2;
    static constexpr int OFF_F32 = 0;
    static constexpr int OFF_H16 = SP_PAIRS * NST * STAGE;
    static constexpr int OFF_RED = OFF_H16 + SP_PAIRS * 2 * SP_XT;
    static constexpr int RED_BYTES = (SP_PAIRS / 2) * SP_NREG * 32 * 4;
    static constexpr int OFF_QF = OFF_RED + RED_BYTES;
    static constexpr int QF_BYTES = 2 * 8 * SP_C * 2 + 32;
    static constexpr int OFF_CSW = OFF_QF + 2 * QF_BYTES;
    static constexpr int OFF_BARS = OFF_CSW + SP_PAIRS * 8 * 4;
    static constexpr int SMEM = OFF_BARS + SP_PAIRS * 6 * 8;          // per pair: full32[2], ready16[2], free16[2]
    static_assert(SMEM <= 232448, "split pass: shared memory budget");
};

}  // namespace

template <int EIN>
__global__ void __launch_bounds__(SP_THREADS, 1) sa_pass1_split_kernel(const SAPassParams p) {
    using Cfg = SplitCfg<EIN>;
    constexpr int C = SP_C, KS = SP_KS, ROWB = SP_ROWB, XT_BYTES = SP_XT, STAGE = Cfg::STAGE, NREG = SP_NREG;
    extern __shared__ __align__(1024) unsigned char smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const bool front = warp < SP_PAIRS;
    const int pr = front ? warp : warp - SP_PAIRS;                      // pair index = pixel-tile stream
    unsigned char* f32s = smem + Cfg::OFF_F32 + pr * Cfg::NST * STAGE;
    unsigned char* h16s = smem + Cfg::OFF_H16 + pr * 2 * XT_BYTES;
    float* red = reinterpret_cast<float*>(smem + Cfg::OFF_RED);
    __half* qf = reinterpret_cast<__half*>(smem + Cfg::OFF_QF);
    float* colsum_w = reinterpret_cast<float*>(smem + Cfg::OFF_CSW);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::OFF_BARS) + pr * 6;
    uint64_t* full32 = bars;          // [2] raw tile landed (TMA complete_tx)
    uint64_t* ready16 = bars + 2;     // [2] fp16 t tile written (front -> back)
    uint64_t* free16 = bars + 4;      // [2] fp16 t tile consumed (back -> front)

    const int N = p.N, K = p.K;
    const int items = p.nframes * p.nchunk;
    const int nbw = p.chunk_px / (16 * SP_PAIRS);     // tiles per pair per item
    const int tiles_chunk = p.chunk_px >> 4;
    const int tiles_frame = p.nchunk * tiles_chunk;
    const int my_items = (items > (int)blockIdx.x) ? (items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
    const uint32_t total_tiles = (uint32_t)my_items * nbw;

    if (front && lane == 0) {
        for (int s = 0; s < 6; ++s) mbar_init(&bars[s], 1);
        fence_mbar_init();
    }
    __syncthreads();
    if (my_items == 0) return;

    // tile n of this pair -> (frame, tile index in the frame, first pixel, valid pixels)
    auto locate = [&](uint32_t n, int& f, int& tb, int& px0, int& nvalid) {
        const int il = n / nbw, j = n % nbw;
        const int item = (int)blockIdx.x + il * (int)gridDim.x;
        f = p.frame0 + item / p.nchunk;
        tb = (item % p.nchunk) * tiles_chunk + pr + SP_PAIRS * j;
        px0 = tb * 16;
        nvalid = N - px0;
        nvalid = nvalid < 0 ? 0 : (nvalid > 16 ? 16 : nvalid);
    };

    if (front) {
        // ======================================= front warps: TMA + LayerNorm =======================================
        const uint64_t pol = l2_policy_evict_first();
        const bool ring_fits_l2 = (size_t)p.xhat_frames * p.n16 * C * 2 <= ((size_t)48 << 20);
        const uint64_t xpol = ring_fits_l2 ? l2_policy_evict_last() : l2_policy_evict_first();
        const int pxi = lane >> 3, ch8 = lane & 7;
        uint32_t nissued = 0;         // TMA loads issued (only tiles with pixels are loaded)
        uint32_t nseq = 0;            // next tile to look at for issuing
        auto issue_next = [&]() {     // issue the next not-yet-issued tile that has pixels (lane 0 only)
            while (nseq < total_tiles) {
                int f, tb, px0, nvalid;
                locate(nseq, f, tb, px0, nvalid);
                ++nseq;
                if (nvalid == 0) continue;
                const int s = nissued & 1;
                const void* src = reinterpret_cast<const unsigned char*>(p.feats) +
                                  ((size_t)f * p.feat_bstride + (size_t)px0 * C) * EIN;
                const uint32_t bytes = (uint32_t)nvalid * C * EIN;
                mbar_arrive_expect_tx(&full32[s], bytes);
                bulk_g2s(f32s + s * STAGE, src, bytes, &full32[s], pol);
                ++nissued;
                return;
            }
        };
        if (lane == 0) { issue_next(); issue_next(); }
        uint32_t m = 0;               // tiles with pixels processed so far (stage / slot / parity counter)
#pragma unroll 1
        for (uint32_t n = 0; n < total_tiles; ++n) {
            int f, tb, px0, nvalid;
            locate(n, f, tb, px0, nvalid);
            if (nvalid == 0) continue;                                  // (warp-uniform) nothing to do, nothing signalled
            const int s = m & 1;
            unsigned char* stg = f32s + s * STAGE;
            mbar_wait(&full32[s], (m >> 1) & 1);
            if (nvalid < 16) {
                // ragged tail: the missing raw rows count as zeros (they normalise to t = 0)
                float4* z = reinterpret_cast<float4*>(stg + (size_t)nvalid * C * EIN);
                for (int i = lane; i < (16 - nvalid) * (C * EIN / 16); i += 32) z[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                __syncwarp();
            }
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
            // the raw stage is in registers: re-arm it with the tile two ahead right away
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) issue_next();

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
            // the fp16 slot: handed back by the back warp (tile m-2) and read by the x^ store of tile m-2
            unsigned char* ht = h16s + s * XT_BYTES;
            if (m >= 2) {
                mbar_wait(&free16[s], ((m >> 1) - 1) & 1);
                if (lane == 0) bulk_wait_read<1>();
                __syncwarp();
            }
#pragma unroll
            for (int rr = 0; rr < 4; ++rr) {
                const int row = 4 * rr + pxi;
                const float mu = sm[rr] * (1.f / C);
                const float var = fmaxf(fmaf(-mu, mu, sq[rr] * (1.f / C)), 0.f);
                const float rstd = rsqrtf(var + SP_LN_EPS);
                const f32x2 r2 = pack2(rstd, rstd);
                const float nb = -mu * rstd;
                const f32x2 nb2 = pack2(nb, nb);
                unsigned char* rowp = ht + row * ROWB + (ch8 & 1) * 8;
#pragma unroll
                for (int i = 0; i < C / 32; ++i) {
                    const f32x2 t0 = fma2(v[rr][2 * i], r2, nb2), t1 = fma2(v[rr][2 * i + 1], r2, nb2);
                    const int chk = ((ch8 >> 1) + 4 * i) ^ (row & 7);
                    uint2 pk; pk.x = pack_h2(lo2(t0), hi2(t0)); pk.y = pack_h2(lo2(t1), hi2(t1));
                    *reinterpret_cast<uint2*>(rowp + chk * 16) = pk;
                }
            }
            fence_proxy_async();          // the t tile (generic writes) precedes the async-proxy read of the store
            __syncwarp();
            if (lane == 0) {
                if (p.xhat != nullptr) {
                    bulk_s2g(p.xhat + ((size_t)(f % p.xhat_frames) * tiles_frame + tb) * (16 * C), ht, XT_BYTES, xpol);
                }
                bulk_commit();
                mbar_arrive(&ready16[s]);
            }
            ++m;
        }
        if (lane == 0) bulk_wait_read<0>();   // no x^ store may still be reading this CTA's shared memory
        return;
    }

    // ========================================= back warps: tensor-core part =========================================
    const int btid = tid - SP_PAIRS * 32;
    const int g = lane >> 2, t4 = lane & 3;
    auto load_qf = [&](int il, int buf, bool async) {
        const int item = (int)blockIdx.x + il * (int)gridDim.x;
        const int f = p.frame0 + item / p.nchunk;
        const uint4* src = reinterpret_cast<const uint4*>(p.qt + (size_t)f * (Cfg::QF_BYTES / 2));
        uint4* dst = reinterpret_cast<uint4*>(qf + (size_t)buf * (Cfg::QF_BYTES / 2));
        for (int i = btid; i < Cfg::QF_BYTES / 16; i += SP_PAIRS * 32) {
            if (async) cp_async16(dst + i, src + i);
            else dst[i] = __ldg(src + i);
        }
        if (async) cp_async_commit();
    };
    auto bsync = [&]() { named_bar_sync(1, SP_PAIRS * 32); };       // the 8 back warps only
    load_qf(0, 0, false);
    bsync();

    uint32_t n = 0, m = 0;
#pragma unroll 1
    for (int il = 0; il < my_items; ++il) {
        const int item = (int)blockIdx.x + il * (int)gridDim.x;
        const int fl = item / p.nchunk, chunk = item % p.nchunk;
        const int f = p.frame0 + fl;
        const __half* qh = qf + (size_t)(il & 1) * (Cfg::QF_BYTES / 2);
        const __half* ql = qh + 8 * C;
        if (il + 1 < my_items) load_qf(il + 1, (il + 1) & 1, true);

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
        const bool s0ok = (2 * t4) < K, s1ok = (2 * t4 + 1) < K;

#pragma unroll 1
        for (int j = 0; j < nbw; ++j, ++n) {
            const int tb = chunk * tiles_chunk + pr + SP_PAIRS * j;
            const int px0 = tb * 16;
            if (px0 >= N) continue;                                   // tile without pixels: the front warp skipped it too
            const int s = m & 1;
            mbar_wait(&ready16[s], (m >> 1) & 1);
            const uint32_t tile_u32 = smem_u32(h16s + s * XT_BYTES);
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
            if (t4 == 3) { pa1 = 1.f / SP_PSCALE; pb1 = 1.f / SP_PSCALE; }   // slot column 7 := 1 (K <= 7): sum_n t[n]
            const __half2 ha = __floats2half2_rn(pa0 * SP_PSCALE, pa1 * SP_PSCALE);
            const __half2 hb = __floats2half2_rn(pb0 * SP_PSCALE, pb1 * SP_PSCALE);
            {
                const float2 fa = __half22float2(ha), fb = __half22float2(hb);
                cs0 += fa.x + fb.x; cs1 += fa.y + fb.y;   // column sums from the ROUNDED values
            }
            const uint32_t b0 = movmatrix_t(*reinterpret_cast<const uint32_t*>(&ha));
            const uint32_t b1 = movmatrix_t(*reinterpret_cast<const uint32_t*>(&hb));
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
            // the slot goes back only after its fragments have been consumed by the MMAs above
            __syncwarp();
            if (lane == 0) mbar_arrive(&free16[s]);
            ++m;
        }

        // ================= item end: reduce the 8 back warps' partials, write them out =================
        float* part = p.partials + ((size_t)f * p.nchunk + chunk) * p.pstride;
        cs0 += __shfl_xor_sync(0xffffffffu, cs0, 4);  cs1 += __shfl_xor_sync(0xffffffffu, cs1, 4);
        cs0 += __shfl_xor_sync(0xffffffffu, cs0, 8);  cs1 += __shfl_xor_sync(0xffffffffu, cs1, 8);
        cs0 += __shfl_xor_sync(0xffffffffu, cs0, 16); cs1 += __shfl_xor_sync(0xffffffffu, cs1, 16);
        if (lane < 4) { colsum_w[pr * 8 + 2 * lane] = cs0; colsum_w[pr * 8 + 2 * lane + 1] = cs1; }
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
            if (pr >= 4) put(pr - 4);
            bsync();
            if (pr < 4) add(pr);
            if (pr >= 2 && pr < 4) put(pr);
            bsync();
            if (pr < 2) add(2 + pr);
            if (pr == 1) put(1);
            bsync();
            if (pr == 0) {
                add(1);
#pragma unroll
                for (int cb = 0; cb < KS; ++cb)
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const int c = 16 * cb + g + 8 * (e >> 1);
                        const int slot = 2 * t4 + (e & 1);
                        part[slot * C + c] = uacc[cb][e];
                        if (slot == 7) part[8 * C + 8 + c] = uacc[cb][e];   // sum_n t[n][c]
                    }
            } else if (pr == 1 && lane < 8) {
                float a = 0.f;
#pragma unroll
                for (int w8 = 0; w8 < SP_PAIRS; ++w8) a += colsum_w[w8 * 8 + lane];
                part[8 * C + lane] = a;
            }
        }
        cp_async_wait_all();
        bsync();
    }
}

template <int EIN>
static cudaError_t split_launch_t(const SAPassParams& p, int sms, cudaStream_t st) {
    using Cfg = SplitCfg<EIN>;
    auto kern = sa_pass1_split_kernel<EIN>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM);
    if (e != cudaSuccess) return e;
    const int items = p.nframes * p.nchunk;
    const int grid = items < sms ? items : sms;
    kern<<<grid, SP_THREADS, Cfg::SMEM, st>>>(p);
    return cudaGetLastError();
}

// first pass on warp pairs; the caller checks that the shape qualifies (sa_pass_split_supported)
cudaError_t sa_pass_split_launch(const SAPassParams& p, int sms, cudaStream_t st) {
    return p.feat_esize == 2 ? split_launch_t<2>(p, sms, st) : split_launch_t<4>(p, sms, st);
}

bool sa_pass_split_supported(const SAPassParams& p, int C) {
    return C == 128 && p.K <= 7 && p.seg_mask == nullptr && p.xhat != nullptr && (p.chunk_px % (16 * SP_PAIRS)) == 0;
}

}  // namespace sfb
