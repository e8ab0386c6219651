// Autoregressive slot-Transformer rollout, engine B: tcgen05 tensor cores + TMEM + TMA.
//
// Same operator as ro_kernel.cu (reference slotformer.py:85-126, single_step_slotformer.py:49-90)
// and the same "one CTA owns one clip for the whole rollout" structure, but every linear layer
// runs on the 5th-generation tensor cores in swap-AB form (the window has few tokens):
//
//     D^T[128 features x Lp tokens] (TMEM fp32) += W[128 x 64] (smem, TMA-fed) * X[Lp x 64]^T (smem)
//
//   warp 8      TMA producer: streams 16 KB weight tiles (128 features x 64 k, pre-swizzled by
//               ro_pack2_kernel) through an mbarrier ring, running ahead across GEMMs and steps;
//   warp 9      one elected thread issues tcgen05.mma (M=128, N=Lp, K=16) and tcgen05.commit;
//               the warp also owns the TMEM allocation (512 columns);
//   warps 0-7   LayerNorm, attention (mma.sync on the small 36x36 score tiles), and the GEMM
//               epilogues: tcgen05.ld TMEM -> registers -> bias / ReLU / residual / positional
//               encoding -> next operand, written straight into the swizzled K-major layout the
//               next tcgen05.mma reads.
// The residual stream stays in shared memory as fp32; activations as fp16 operands.
#include "ro_attn.cuh"
#include "umma.cuh"
#include "ro_kernel.h"

namespace sfb {

static constexpr int RU_TILE_BYTES = 16384;                 // one 128 x 64 weight tile
static constexpr int RU_HPAD = 4;                         // fp32 pad of a residual-stream row (bank spread for 4-lane row groups)
static constexpr int RU_MAX_STAGE_TILES = 2;                // k-adjacent tiles per ring stage: 1 or 2 (plan)
static constexpr int RU_TILE_HALVES = 8192;
static constexpr int RU_SYNC_THREADS = RO_THREADS + 32;   // compute warps + MMA warp
static constexpr int RU_THREADS = RO_THREADS + 128;   // + one warpgroup: TMA producer, MMA issuer, two idle warps
static constexpr int RU_REGS_COMPUTE = 232;          // setmaxnreg: the epilogue / attention warps take what the
static constexpr int RU_REGS_SERVICE = 40;           // service warpgroup gives up (8*32*232 + 4*32*40 = 64512)

struct UOp {
    const __half* base;   // packed matrix (ro_pack2 layout)
    int kpt;              // 64-wide k blocks per feature tile of the whole matrix
    int tile0, ntile;     // 128-feature tiles
    int kb0, nkb;         // k-block range
};

// barrier words (uint64 each) inside the `bars` region
static constexpr int RU_BAR_FULL = 0;     // [8] weight stage filled (TMA complete_tx)
static constexpr int RU_BAR_EMPTY = 8;    // [8] weight stage consumed (tcgen05.commit)
static constexpr int RU_BAR_ACC = 16;     // [9] accumulator tile t complete; [8] = second GEMM of the fused FFN
static constexpr int RU_BAR_RDY = 25;     // [8] hidden k-block pair j written by the epilogue warps
static constexpr int RU_BAR_CHUNK = 33;   // FFN chunk boundary: the chunk's W2 MMAs have read the hidden tile
static constexpr int RU_BAR_PAR = 34;     // [2] per-layer parameter block landed
static constexpr int RU_BAR_TMEM = 36;    // TMEM base address slot
static constexpr int RU_BAR_WORDS = 40;
static constexpr int RU_ACC2 = 8;

// extras of an fp16-epilogue GEMM call: absolute first feature of op tile 0 (bias / column index), first token row
// of the N range, first accumulator barrier and first TMEM column to use
struct H16Ext { int feat0, row0, tbase; uint32_t dcol0; };

struct BRing {
    unsigned char* stages;
    uint64_t* full;
    uint64_t* empty;
    int nstage;
    int stage_tiles;      // weight tiles (16 KB) per stage
};

// arguments of the fused feed-forward block  h += W2 relu(W1 x^ + b1) + b2  (one hidden chunk)
struct FfnArgs {
    const __half *w1, *w2;
    const float *b1, *b2;      // smem copies (b1 indexed by absolute hidden feature)
    int d, F, f0, fcw;         // model width, hidden width, chunk start / width
    int Lp, kbb;               // tokens (MMA N) of this call, bytes between k blocks
    int row0;                  // first token row of this call (last-layer pruning: only the tail rows)
    uint32_t x_u32, y_u32;     // operand buffers at row 0
    unsigned char* yb;
    float* h;
    uint32_t acc2_col;         // TMEM column of the W2 accumulator
    bool first, last;          // first / last hidden chunk of the layer
};

// one lane polls, the rest of the warp sleeps at the warp barrier (8 pollers instead of 256)
__device__ __forceinline__ void mbar_wait_warp(uint64_t* bar, uint32_t parity, int lane) {
    if (lane == 0) { while (!mbar_try_wait(bar, parity)) { } }
    __syncwarp();      // lane 0's acquire + the warp barrier order the other lanes' accesses behind the phase
}

// fine-grained per-role trace (clock64 of this SM) of one layer, for the debug timeline only
// (compiled in only with -DSFB_FINE_PROF: the trace state costs registers the attention phase needs)
#ifdef SFB_FINE_PROF
struct FineProf {
    unsigned long long* buf;   // nullptr = off
    int n, cap;
    bool on;
    __device__ __forceinline__ void mark(int tag) {
        if (buf != nullptr && on && n < cap) buf[n++] = ((unsigned long long)tag << 48) | ((unsigned long long)clock64() & 0xFFFFFFFFFFFFull);
    }
    __device__ __forceinline__ void enable(bool v) { on = v; }
};
#else
struct FineProf {
    __device__ __forceinline__ FineProf(unsigned long long*, int, int, bool) {}
    __device__ __forceinline__ void mark(int) {}
    __device__ __forceinline__ void enable(bool) {}
};
#endif

struct BProducer {
    static constexpr bool kCompute = false;
    BRing ring;
    uint32_t pidx;            // next stage
    uint32_t phase;           // its fill parity
    uint64_t pol;
    int dbg;
    FineProf fp;
    __device__ __forceinline__ void emit(const UOp& op) {
        for (int t = 0; t < op.ntile; ++t)
            for (int kb = 0; kb < op.nkb; kb += ring.stage_tiles) {
                const int nk = (op.nkb - kb) < ring.stage_tiles ? (op.nkb - kb) : ring.stage_tiles;
                // stage index / phase kept incrementally (pidx % nstage and pidx / nstage with a run-time nstage
                // are ~40-cycle integer sequences on the handshake path)
                const int s = (int)pidx;
                const uint32_t par = phase;
                if (++pidx == (uint32_t)ring.nstage) { pidx = 0; phase ^= 1u; }
                // spin (no back-off): a sleeping producer adds its wake-up latency to every refill of the ring
                fp.mark(1);
                mbar_wait(&ring.empty[s], par ^ 1u);
                fp.mark(2);
                if (dbg & 1) { mbar_arrive(&ring.full[s]); continue; }
                mbar_arrive_expect_tx(&ring.full[s], nk * RU_TILE_BYTES);
                bulk_g2s(ring.stages + (size_t)s * ring.stage_tiles * RU_TILE_BYTES,
                         op.base + ((size_t)(op.tile0 + t) * op.kpt + op.kb0 + kb) * RU_TILE_HALVES,
                         nk * RU_TILE_BYTES, &ring.full[s], pol);
            }
    }
    template <int HF = 0, class Pre, class Epi>
    __device__ __forceinline__ void gemm(const UOp& op, uint32_t, int, int, Pre, Epi) { emit(op); }
    // three-term product (W_hi X_hi + W_hi X_lo + W_lo X_hi): the hi tiles pass through the ring twice
    template <int HF = 0, class Pre, class Epi>
    __device__ __forceinline__ void gemm3(const UOp& hi, const UOp& lo, uint32_t, uint32_t, int, int, Pre, Epi) {
        emit(hi); emit(hi); emit(lo);
    }
    template <int HMAX>
    __device__ __forceinline__ void gemm_h16(const UOp& op, uint32_t, int, int, const float*, float, int, unsigned char*, H16Ext) { emit(op); }
    // weight order of the fused FFN: every W1 tile of the chunk, then the W2 k-block pairs in tile order
    template <int HMAX>
    __device__ __forceinline__ void ffn(const FfnArgs& a) {
        const int nt = a.fcw >> 7, dt = a.d >> 7, kpd = a.d >> 6;
        emit(UOp{a.w1, kpd, a.f0 >> 7, nt, 0, kpd});
        for (int j = 0; j < nt; ++j) emit(UOp{a.w2, a.F >> 6, 0, dt, (a.f0 >> 6) + 2 * j, 2});
    }
    __device__ __forceinline__ void sync() {}
};

struct BMma {
    static constexpr bool kCompute = false;
    BRing ring;
    uint32_t pidx;            // next stage
    uint32_t phase;           // its fill parity
    uint32_t tmem;
    uint64_t* bars;
    uint32_t rdypar;          // bit j = parity of the next completion of rdy[j]
    int lane;
    int dbg;
    FineProf fp;
    // MMAs of `op` into TMEM columns dcol + t * ntok (tile t); b_u32: smem address of the token operand's
    // k-block 0 of this op, kblock_bytes apart per 64 k.  acc_first: the first k-step accumulates too.
    __device__ __forceinline__ void issue(const UOp& op, uint32_t b_u32, int kblock_bytes, int ntok, uint32_t dcol,
                                          bool acc_first, bool commit_tiles, int tbase = 0) {
        const uint32_t idesc = umma_idesc_f16(128, ntok);
        for (int t = 0; t < op.ntile; ++t) {
            for (int kb = 0; kb < op.nkb; kb += ring.stage_tiles) {
                const int nk = (op.nkb - kb) < ring.stage_tiles ? (op.nkb - kb) : ring.stage_tiles;
                const int s = (int)pidx;
                const uint32_t par = phase;
                if (++pidx == (uint32_t)ring.nstage) { pidx = 0; phase ^= 1u; }
                fp.mark(1);
                mbar_wait(&ring.full[s], par);
                fp.mark(2);
                tcgen05_fence_after();
                // descriptors are built once per stage; a K step of 16 (32 B) or the next k tile only
                // advances the 16-byte-granular start-address field
                const uint64_t da0 = umma_smem_desc(smem_u32(ring.stages + (size_t)s * ring.stage_tiles * RU_TILE_BYTES));
                const uint64_t db0 = umma_smem_desc(b_u32 + kb * kblock_bytes);
                const uint32_t dtm = tmem + dcol + (uint32_t)(t * ntok);
                if (elect_one()) {
                    if (!(dbg & 2)) {
#pragma unroll
                        for (int kk = 0; kk < RU_MAX_STAGE_TILES; ++kk) {
                            if (kk < nk) {
                                const uint64_t dbk = db0 + (uint64_t)((kk * kblock_bytes) >> 4);
#pragma unroll
                                for (int k4 = 0; k4 < 4; ++k4)
                                    umma_f16(dtm, da0 + (uint64_t)((kk * RU_TILE_BYTES + k4 * 32) >> 4),
                                             dbk + (uint64_t)((k4 * 32) >> 4), idesc, acc_first || (kb | kk | k4) != 0);
                            }
                        }
                    }
                    umma_commit(&ring.empty[s]);           // stage is free once these MMAs have read it
                }
                __syncwarp();
                fp.mark(3);
            }
            if (commit_tiles) commit(&bars[RU_BAR_ACC + tbase + t]);   // this tile's accumulator is complete
        }
    }
    __device__ __forceinline__ void commit(uint64_t* bar) {
        if (elect_one()) umma_commit(bar);
        __syncwarp();
    }
    template <int HF = 0, class Pre, class Epi>
    __device__ __forceinline__ void gemm(const UOp& op, uint32_t b_u32, int kblock_bytes, int ntok, Pre, Epi) {
        issue(op, b_u32, kblock_bytes, ntok, 0u, false, true);      // whole warp, uniform; one elected lane issues
    }
    template <int HF = 0, class Pre, class Epi>
    __device__ __forceinline__ void gemm3(const UOp& hi, const UOp& lo, uint32_t bhi_u32, uint32_t blo_u32, int kblock_bytes,
                                          int ntok, Pre, Epi) {
        issue(hi, bhi_u32, kblock_bytes, ntok, 0u, false, false);
        issue(hi, blo_u32, kblock_bytes, ntok, 0u, true, false);
        issue(lo, bhi_u32, kblock_bytes, ntok, 0u, true, true);     // tile t's accumulator is complete after its third term
    }
    template <int HMAX>
    __device__ __forceinline__ void gemm_h16(const UOp& op, uint32_t b_u32, int kblock_bytes, int ntok, const float*, float,
                                             int, unsigned char*, H16Ext x) {
        issue(op, b_u32 + (uint32_t)(x.row0 * 128), kblock_bytes, ntok, x.dcol0, false, true, x.tbase);
    }
    __device__ __forceinline__ void ffn2_part(const FfnArgs& a, int j) {
        fp.mark(4);
        mbar_wait(&bars[RU_BAR_RDY + j], (rdypar >> j) & 1u);
        fp.mark(5);
        rdypar ^= 1u << j;
        tcgen05_fence_after();
        issue(UOp{a.w2, a.F >> 6, 0, a.d >> 7, (a.f0 >> 6) + 2 * j, 2}, a.y_u32 + (uint32_t)(2 * j * a.kbb + a.row0 * 128), a.kbb, a.Lp,
              a.acc2_col, !(a.first && j == 0), false);
    }
    // all W1 tiles first (they only need x^), then each W2 k-block pair as soon as the epilogue warps have
    // written that hidden tile: the W2 MMAs of tile j overlap the epilogue of tile j+1
    template <int HMAX>
    __device__ __forceinline__ void ffn(const FfnArgs& a) {
        const int nt = a.fcw >> 7, kpd = a.d >> 6;
        issue(UOp{a.w1, kpd, a.f0 >> 7, nt, 0, kpd}, a.x_u32 + (uint32_t)(a.row0 * 128), a.kbb, a.Lp, 0u, false, true);
        for (int j = 0; j < nt; ++j) ffn2_part(a, j);
        commit(a.last ? &bars[RU_BAR_ACC + RU_ACC2] : &bars[RU_BAR_CHUNK]);
    }
    __device__ __forceinline__ void sync() {
        named_bar_sync(1, RU_SYNC_THREADS);
        tcgen05_fence_after();
    }
};

struct BCompute {
    static constexpr bool kCompute = true;
    uint32_t tmem;
    uint64_t* bars;
    uint32_t accpar;          // bit t = parity of the next completion of acc[t]
    uint32_t chunkpar;
    int warp, lane;
    int dbg;
    FineProf fp;
    __device__ __forceinline__ void acc_wait(int t) {
        fp.mark(1);
        mbar_wait_warp(&bars[RU_BAR_ACC + t], (accpar >> t) & 1u, lane);
        fp.mark(2);
        accpar ^= 1u << t;
        tcgen05_fence_after();
    }
    // epi(feature, token8, values[8], pre(feature)): 8 consecutive tokens (token8 % 8 == 0) of one feature;
    // features of the warp's TMEM lane quadrant, one half of the tokens (warps w and w+4 share lanes)
    // HFIX > 0: tokens per warp known at compile time (the usual full window): the guards below fold away
    template <int HFIX = 0, class Pre, class Epi>
    __device__ __forceinline__ void epi_tile(int t, uint32_t dcol, int ntok, Pre pre, Epi epi) {
        if (dbg & 4) return;
        const int q = warp & 3, half = HFIX > 0 ? HFIX : (ntok >> 1), t0 = (warp >> 2) * half;
        const int f = t * 128 + 32 * q + lane;
        const float pv = pre(f);
        const uint32_t ta = tmem + ((uint32_t)(32 * q) << 16) + dcol + (uint32_t)t0;
        float v[4][8];                               // up to 32 tokens per warp: every load first, one wait
#pragma unroll
        for (int c = 0; c < 4; ++c)
            if (8 * c < half) tmem_ld8(ta + 8 * c, v[c]);
        tmem_ld_wait();
#pragma unroll
        for (int c = 0; c < 4; ++c)
            if (8 * c < half) epi(f, t0 + 8 * c, v[c], pv);      // 8 consecutive tokens, base a multiple of 8
    }
    // fp16 epilogue of one 128-feature tile: (acc + bias) [* qscale for features < nscale] [relu] -> operand
    // buffer `yb` (token rows, K-major swizzled).  16x256b TMEM loads give the stmatrix.trans fragment, so a
    // 16-feature x 16-token block is one store instruction instead of 8 scattered 2-byte stores per thread.
    template <int HMAX, bool FULL = false>
    __device__ __forceinline__ void epi_tile_h16(int t, uint32_t dcol, int ntok, const float* bias, bool relu,
                                                 float qscale, int nscale, unsigned char* yb, int kbb, int feat0 = 0,
                                                 int row0 = 0) {
        if (dbg & 4) return;
        // HMAX >= half: tokens per warp; FULL: half == HMAX (the usual full window), guards fold at compile time
        const int q = warp & 3, half = FULL ? HMAX : (ntok >> 1), t0 = (warp >> 2) * half;
        const uint32_t yb_u32 = smem_u32(yb);
        float v[2][HMAX / 2];        // [16-feature group][4 values per 8-token block]
#pragma unroll
        for (int g = 0; g < 2; ++g) {
            const uint32_t ta = tmem + ((uint32_t)(32 * q + 16 * g) << 16) + dcol + (uint32_t)t0;
#pragma unroll
            for (int c0 = 0; c0 < HMAX; c0 += 16) {
                if (c0 + 16 <= HMAX && c0 + 16 <= half) {
                    float w[8];
                    tmem_ld_16x256b_x2(ta + c0, w);
#pragma unroll
                    for (int i = 0; i < 8; ++i) v[g][(c0 >> 1) + i] = w[i];
                } else if (c0 < half) {
                    float w[4];
                    tmem_ld_16x256b_x1(ta + c0, w);
#pragma unroll
                    for (int i = 0; i < 4; ++i) v[g][(c0 >> 1) + i] = w[i];
                }
            }
        }
        fp.mark(20);
        tmem_ld_wait();
        fp.mark(21);
        const int mi = lane >> 3, r = lane & 7;    // lane j supplies the row address of matrix j/8, row j%8
#pragma unroll
        for (int g = 0; g < 2; ++g) {
            const int fb = feat0 + t * 128 + 32 * q + 16 * g;  // first (absolute) feature of the group
            const float b0 = bias[fb + (lane >> 2)], b8 = bias[fb + 8 + (lane >> 2)];
            const float sc = (fb < nscale) ? qscale : 1.f;
            const int f8 = fb + (mi & 1) * 8;
            const uint32_t abase = yb_u32 + (uint32_t)((f8 >> 6) * kbb + (((f8 & 63) >> 3) << 4));
#pragma unroll
            for (int c0 = 0; c0 < HMAX; c0 += 16) {
                if (c0 < half) {
                    const bool full = (c0 + 16 <= HMAX) && (c0 + 16 <= half);
                    uint32_t m[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        if (j < 2 || c0 + 16 <= HMAX) {
                            const float bi = (j & 1) ? b8 : b0;
                            float x0 = (v[g][(c0 >> 1) + 2 * j] + bi) * sc, x1 = (v[g][(c0 >> 1) + 2 * j + 1] + bi) * sc;
                            if (relu) { x0 = fmaxf(x0, 0.f); x1 = fmaxf(x1, 0.f); }
                            m[j] = pack_h2(x0, x1);
                        } else {
                            m[j] = 0u;
                        }
                    }
                    const int tok = row0 + t0 + c0 + (full ? (mi >> 1) * 8 : 0) + r;
                    const uint32_t addr = abase + (uint32_t)(tok * 128) ;
                    const uint32_t swz = (uint32_t)((tok & 7) << 4);
                    if (full) stsm_x4_t(addr ^ swz, m[0], m[1], m[2], m[3]);
                    else stsm_x2_t(addr ^ swz, m[0], m[1]);
                }
            }
        }
    }
    // QKV-style GEMM: every tile goes through the fp16 epilogue
    template <int HMAX>
    __device__ __forceinline__ void gemm_h16(const UOp& op, uint32_t, int kbb, int ntok, const float* bias, float qscale,
                                             int nscale, unsigned char* yb, H16Ext x) {
        const bool full = ntok == 2 * HMAX;
        for (int t = 0; t < op.ntile; ++t) {
            acc_wait(x.tbase + t);
            if (full) epi_tile_h16<HMAX, true>(t, x.dcol0 + (uint32_t)(t * ntok), ntok, bias, false, qscale, nscale, yb, kbb, x.feat0, x.row0);
            else epi_tile_h16<HMAX, false>(t, x.dcol0 + (uint32_t)(t * ntok), ntok, bias, false, qscale, nscale, yb, kbb, x.feat0, x.row0);
        }
        tcgen05_fence_before();
    }
    // HF > 0: tokens per warp of the usual full window; that case runs the epilogue whose guards are constants
    template <int HF = 0, class Pre, class Epi>
    __device__ __forceinline__ void gemm(const UOp& op, uint32_t, int, int ntok, Pre pre, Epi epi) {
        const bool full = HF > 0 && ntok == 2 * HF;
        for (int t = 0; t < op.ntile; ++t) {
            acc_wait(t);
            if (full) epi_tile<HF>(t, (uint32_t)(t * ntok), ntok, pre, epi);   // overlaps the MMAs of tile t+1
            else epi_tile<0>(t, (uint32_t)(t * ntok), ntok, pre, epi);
        }
        tcgen05_fence_before();
    }
    template <int HF = 0, class Pre, class Epi>
    __device__ __forceinline__ void gemm3(const UOp& hi, const UOp&, uint32_t, uint32_t, int, int ntok, Pre pre, Epi epi) {
        gemm<HF>(hi, 0u, 0, ntok, pre, epi);
    }
    template <int HMAX>
    __device__ __forceinline__ void ffn(const FfnArgs& a) {
        const int nt = a.fcw >> 7, dt = a.d >> 7;
        if (!a.first) {       // the previous chunk's W2 MMAs must have read the hidden tile before it is overwritten
            mbar_wait_warp(&bars[RU_BAR_CHUNK], chunkpar & 1u, lane);
            chunkpar ^= 1u;
        }
        unsigned char* yb = a.yb;
        const int kbb = a.kbb;
        const bool full = a.Lp == 2 * HMAX;
        for (int t = 0; t < nt; ++t) {
            acc_wait(t);
            if (full) epi_tile_h16<HMAX, true>(t, (uint32_t)(t * a.Lp), a.Lp, a.b1 + a.f0, true, 1.f, 0, yb, kbb, 0, a.row0);
            else epi_tile_h16<HMAX, false>(t, (uint32_t)(t * a.Lp), a.Lp, a.b1 + a.f0, true, 1.f, 0, yb, kbb, 0, a.row0);
            fp.mark(22);
            tcgen05_fence_before();
            fp.mark(23);
            fence_proxy_async();                     // hidden k-blocks 2t, 2t+1 -> tensor-core reads
            fp.mark(24);
            __syncwarp();
            if (lane == 0) mbar_arrive(&bars[RU_BAR_RDY + t]);
            fp.mark(3);
        }
        if (a.last) {
            acc_wait(RU_ACC2);
            float* h = a.h;
            const int d = a.d + RU_HPAD;
            for (int t2 = 0; t2 < dt; ++t2)
            {
                auto pre2 = [&](int f) { return a.b2[f]; };
                auto epi2 = [&](int f, int t8, const float (&v)[8], float bi) {
                    float* hf = h + (a.row0 + t8) * d + f;   // d: padded row stride
#pragma unroll
                    for (int i = 0; i < 8; ++i) hf[i * d] += v[i] + bi;
                };
                if (full) epi_tile<HMAX>(t2, a.acc2_col + (uint32_t)(t2 * a.Lp), a.Lp, pre2, epi2);
                else epi_tile<0>(t2, a.acc2_col + (uint32_t)(t2 * a.Lp), a.Lp, pre2, epi2);
            }
            tcgen05_fence_before();
        }
    }
    __device__ __forceinline__ void sync() {
        fence_proxy_async();   // operand tiles written by these threads -> tensor-core reads
        named_bar_sync(1, RU_SYNC_THREADS);
    }
};

template <int DMODEL, int DH, int NKB, class Role>
__device__ __forceinline__ void run_rollout_b(Role& R, const ROParams& p, float* h, unsigned char* xb,
                                              unsigned char* yb, float* par, uint64_t* bars, int tid, int warp, int lane) {
    const int K = p.K, Ds = p.Ds, F = p.F, FC = p.fc;
    const int LpMax = (p.lmax + 15) & ~15;
    const int kbb = LpMax * 128;                       // bytes between 64-wide k blocks of X / Y
    const uint32_t x_u32 = smem_u32(xb), y_u32 = smem_u32(yb);
    const float sm_scale_log2 = rsqrtf((float)DH) * 1.4426950408889634f;
    auto addr_s = [=](int r, int c) { return swz_off(r, c, kbb); };
    const int PF = p.par_floats;
    // per-layer parameter block (biases + LayerNorm affine, packed by sfb_rollout_prepare): one bulk copy
    auto load_params = [&](int layer, int buf) {
        if (Role::kCompute && tid == 0) {
            mbar_arrive_expect_tx(&bars[RU_BAR_PAR + buf], (uint32_t)PF * 4u);
            bulk_g2s(par + (size_t)buf * PF, p.par_g + (size_t)layer * PF, (uint32_t)PF * 4u, &bars[RU_BAR_PAR + buf],
                     l2_policy_evict_last());
        }
    };
    int pidx_prof = 0;
    const bool do_prof = Role::kCompute && (p.prof != nullptr) && blockIdx.x == 0 && tid == 0;
    auto stamp = [&]() { if (do_prof && pidx_prof < p.prof_cap) p.prof[pidx_prof++] = globaltimer_ns(); };
    uint32_t lcount = 0;
    const int my_clips = ((int)blockIdx.x < p.B) ? (p.B - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
    const uint32_t total_layers = (uint32_t)my_clips * p.pred_len * p.layers;
    load_params(0, 0);
    R.sync();

    for (int b = blockIdx.x; b < p.B; b += gridDim.x) {
        const float* hist = p.hist + (size_t)b * p.hist_tokens * Ds;
        float* pred = p.pred + (size_t)b * p.pred_len * K * Ds;
        for (int step = 0; step < p.pred_len; ++step) {
            const int total = p.hist_tokens + step * K;
            int L, base, pe0;
            if (p.mode == 0) { L = p.hist_tokens; base = step * K; pe0 = 0; }
            else { L = total < p.cond_tokens ? total : p.cond_tokens; base = total - L; pe0 = p.pe_tokens - L; }
            const int Lp = (L + 15) & ~15, nmb = Lp >> 4, nkb = (L + 7) >> 3;   // key blocks holding a valid key

            stamp();   // 0 step start
            if (Role::kCompute) {
                for (int i = tid; i < Lp * (Ds / 4); i += RO_THREADS) {
                    const int r = i / (Ds / 4), c4 = (i % (Ds / 4)) * 4;
                    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (r < L) {
                        const int a = base + r;
                        const float* src = (a < p.hist_tokens) ? hist + (size_t)a * Ds
                                                               : pred + (size_t)(a - p.hist_tokens) * Ds;
                        v = *reinterpret_cast<const float4*>(src + c4);
                    }
                    // the slots enter as hi + lo fp16 pairs (lo in the Y buffer, free until the QKV epilogue): in_proj
                    // and out_proj are three-term products, their operand rounding would otherwise be a third of the
                    // step's error budget (scripts/ro_error_budget.py)
                    uint2 pk, pl;
                    pk.x = pack_h2(v.x, v.y); pk.y = pack_h2(v.z, v.w);
                    const float2 h0 = __half22float2(*reinterpret_cast<const __half2*>(&pk.x));
                    const float2 h1 = __half22float2(*reinterpret_cast<const __half2*>(&pk.y));
                    pl.x = pack_h2(v.x - h0.x, v.y - h0.y); pl.y = pack_h2(v.z - h1.x, v.w - h1.y);
                    *reinterpret_cast<uint2*>(xb + addr_s(r, c4)) = pk;
                    *reinterpret_cast<uint2*>(yb + addr_s(r, c4)) = pl;
                }
            }
            R.sync();
            // ---- in_proj + positional encoding -> h ----
            {
                const UOp op{p.w_in, Ds >> 6, 0, DMODEL >> 7, 0, Ds >> 6};
                const UOp op_lo{p.w_in_lo, Ds >> 6, 0, DMODEL >> 7, 0, Ds >> 6};
                R.template gemm3<NKB * 4>(op, op_lo, x_u32, y_u32, kbb, Lp,
                       [&](int f) { return __ldg(p.b_in + f); },
                       [&](int f, int t8, const float (&v)[8], float bi) {
#pragma unroll
                           for (int i = 0; i < 8; ++i) {
                               const int t = t8 + i;
                               h[t * (DMODEL + RU_HPAD) + f] = (t < L) ? v[i] + bi + __ldg(p.pe + (size_t)(pe0 + t) * DMODEL + f) : 0.f;
                           }
                       });
            }
            R.sync();
            stamp();   // in_proj done

            for (int layer = 0; layer < p.layers; ++layer) {
                const ROLayer& ly = p.layer[layer];
                R.fp.enable(step == 1 && layer == 1);
                R.fp.mark(9);
                const float* pb = par + (size_t)(lcount & 1) * PF;
                const float* s_bqkv = pb;
                const float* s_bo = pb + 3 * DMODEL;
                const float* s_b1 = pb + 4 * DMODEL;
                const float* s_b2 = pb + 4 * DMODEL + F;
                const float* l1w = pb + 5 * DMODEL + F;
                const float* l1b = l1w + DMODEL;
                const float* l2w = l1w + 2 * DMODEL;
                const float* l2b = l1w + 3 * DMODEL;
                // Last layer: only the last K tokens feed out_proj (slotformer.py:121), so queries, attention output,
                // out-proj, LN2 and the feed-forward block are needed for those rows only; keys / values still for every
                // token.  [rq0, rq0 + nq): the 16-aligned-size, 8-aligned-start row range that covers [L - K, L).
                int rq0 = 0, nq = Lp;
                if (layer == p.layers - 1 && !(p.dbg & 128)) {
                    int r0 = (L - K) & ~7;
                    int n = (L - r0 + 15) & ~15;
                    if (r0 + n > Lp) r0 = Lp - n;
                    if (n < Lp) { rq0 = r0; nq = n; }
                }
                const bool pruned = nq < Lp;
                if (Role::kCompute) {
                    // next layer's parameters: its buffer was last read before the barrier that ended the previous layer
                    if (lcount + 1 < total_layers) load_params((layer + 1) % p.layers, (lcount + 1) & 1);
                    mbar_wait_warp(&bars[RU_BAR_PAR + (lcount & 1)], (lcount >> 1) & 1u, lane);
                    if (!(p.dbg & 64)) ln_to_half_w<DMODEL, RU_HPAD, NKB / 2>(h, xb, addr_s, L, Lp, l1w, l1b, warp, lane);
                }
                R.sync();
                stamp();   // LN1
                R.fp.mark(11);
                // ---- packed q | k | v projection -> Y[token][3d]; q pre-scaled by log2(e)/sqrt(dh) ----
                // full layers: one call over q | k | v with every token.  Pruned last layer: k | v of every token, then
                // q of the tail rows (own accumulator barriers and TMEM columns).  One call site (a 2-trip loop) so that the
                // epilogue is instantiated once.
#pragma unroll 1
                for (int part = 0; part < (pruned ? 2 : 1); ++part) {
                    const bool qpart = pruned && part == 1;
                    const int tile0 = (pruned && part == 0) ? (DMODEL >> 7) : 0;
                    const int ntile = !pruned ? ((3 * DMODEL) >> 7) : (qpart ? (DMODEL >> 7) : ((2 * DMODEL) >> 7));
                    const UOp op{ly.wqkv, DMODEL >> 6, tile0, ntile, 0, DMODEL >> 6};
                    R.template gemm_h16<NKB * 4>(op, x_u32, kbb, qpart ? nq : Lp, s_bqkv, sm_scale_log2, DMODEL, yb,
                                                 H16Ext{tile0 * 128, qpart ? rq0 : 0, qpart ? 4 : 0,
                                                        qpart ? (uint32_t)(((2 * DMODEL) >> 7) * Lp) : 0u});
                }
                R.sync();
                stamp();   // qkv
                R.fp.mark(12);
                if (Role::kCompute && !(p.dbg & 32)) {
                    if (NKB <= 6) {
                        // (the pruned last layer has nq / 16 query blocks starting at row rq0; same call, same code)
                        constexpr int AKB = NKB <= 6 ? NKB : 2, AMB = NKB <= 6 ? NKB / 2 : 1;
                        // the usual full layer (every query block, all but the last key block valid: L = 36 in a 48-row
                        // window, L = 42 ...) runs the instantiation whose guards are compile-time constants
                        const bool usual = (nq >> 4) == AMB && nkb == AKB - 1;
                        for (int hh = warp; hh < p.heads; hh += RO_WARPS) {
                            if (usual)
                                attn_head<DH, AKB, AMB, AKB - 1, AMB>(
                                    yb, addr_s, AMB, hh * DH, DMODEL + hh * DH, 2 * DMODEL + hh * DH, L, AKB - 1, 1.f, lane, rq0);
                            else
                                attn_head<DH, AKB, AMB>(
                                    yb, addr_s, nq >> 4, hh * DH, DMODEL + hh * DH, 2 * DMODEL + hh * DH, L, nkb, 1.f, lane, rq0);
                        }
                    } else {
                        const int nblk = nq >> 4;
                        for (int item = warp; item < p.heads * nblk; item += RO_WARPS) {
                            const int hh = item / nblk, qb = item % nblk;
                            attn_rows<DH, NKB>(yb, addr_s, rq0 + 16 * qb, hh * DH, DMODEL + hh * DH, 2 * DMODEL + hh * DH, L, nkb,
                                               1.f, lane);
                        }
                    }
                }
                R.sync();
                stamp();   // attn
                R.fp.mark(13);
                // ---- h += O Wo^T + bo ----
                {
                    const UOp op{ly.wo, DMODEL >> 6, 0, DMODEL >> 7, 0, DMODEL >> 6};
                    R.template gemm<NKB * 4>(op, y_u32 + (uint32_t)(rq0 * 128), kbb, nq,
                           [&](int f) { return s_bo[f]; },
                           [&](int f, int t8, const float (&v)[8], float bi) {
                               float* hf = h + (rq0 + t8) * (DMODEL + RU_HPAD) + f;
#pragma unroll
                               for (int i = 0; i < 8; ++i) hf[i * (DMODEL + RU_HPAD)] += v[i] + bi;
                           });
                }
                R.sync();
                stamp();   // outproj
                R.fp.mark(14);
                R.fp.mark(30);
                if (Role::kCompute && !(p.dbg & 64))
                    ln_to_half_w<DMODEL, RU_HPAD, NKB / 2>(h, xb, addr_s, L, rq0 + nq, l2w, l2b, warp, lane, rq0);
                R.fp.mark(31);
                R.sync();
                R.fp.mark(32);
                stamp();   // LN2
                R.fp.mark(15);
                // ---- h += W2 relu(W1 y + b1) + b2: fused, the W2 MMAs of hidden tile t start as soon as its
                //      epilogue has written the tile; FC hidden features per chunk ----
                for (int f0 = 0; f0 < F; f0 += FC) {
                    const int fcw = (F - f0) < FC ? (F - f0) : FC;
                    const FfnArgs fa{ly.w1, ly.w2, s_b1, s_b2, DMODEL, F, f0, fcw, nq, kbb, rq0, x_u32, y_u32, yb, h,
                                     (uint32_t)((FC >> 7) * LpMax), f0 == 0, f0 + FC >= F};
                    R.template ffn<NKB * 4>(fa);
                }
                R.sync();
                stamp();   // ffn
                R.fp.mark(10);
                R.fp.enable(false);
                ++lcount;
            }

            // ---- out_proj on the last K tokens -> pred_out[b, step] ----
            if (Role::kCompute) {
                for (int i = tid; i < 16 * DMODEL; i += RO_THREADS) {
                    const int r = i / DMODEL, c = i % DMODEL;
                    const float v = r < K ? h[(L - K + r) * (DMODEL + RU_HPAD) + c] : 0.f;
                    const __half hi = __float2half_rn(v);
                    *reinterpret_cast<__half*>(xb + addr_s(r, c)) = hi;
                    *reinterpret_cast<__half*>(yb + addr_s(r, c)) = __float2half_rn(v - __half2float(hi));
                }
            }
            R.sync();
            float* dst = pred + (size_t)step * K * Ds;
            {
                const UOp op{p.w_out, DMODEL >> 6, 0, (Ds + 127) >> 7, 0, DMODEL >> 6};
                const UOp op_lo{p.w_out_lo, DMODEL >> 6, 0, (Ds + 127) >> 7, 0, DMODEL >> 6};
                R.gemm3(op, op_lo, x_u32, y_u32, kbb, 16,
                       [&](int f) { return f < Ds ? __ldg(p.b_out + f) : 0.f; },
                       [&](int f, int t8, const float (&v)[8], float bi) {
#pragma unroll
                           for (int i = 0; i < 8; ++i)
                               if (t8 + i < K && f < Ds) dst[(size_t)(t8 + i) * Ds + f] = v[i] + bi;
                       });
            }
            R.sync();   // pred_out[step] visible to this CTA's next window load
        }
    }
}

template <int DMODEL, int DH, int NKB>
__global__ void __launch_bounds__(RU_THREADS, 1) ro_umma_forward_kernel(const ROParams p) {
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char* xb = smem + p.off_a;
    unsigned char* yb = smem + p.off_b;
    float* h = reinterpret_cast<float*>(smem + p.off_h);
    float* par = reinterpret_cast<float*>(smem + p.off_par);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + p.off_bars);
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + RU_BAR_TMEM);
    BRing ring{smem + p.off_ring, bars + RU_BAR_FULL, bars + RU_BAR_EMPTY, p.nstage, p.stage_tiles};
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (tid == 0) {
        for (int s = 0; s < p.nstage; ++s) { mbar_init(&ring.full[s], 1); mbar_init(&ring.empty[s], 1); }
        for (int i = 0; i < 9; ++i) mbar_init(&bars[RU_BAR_ACC + i], 1);
        for (int i = 0; i < 8; ++i) mbar_init(&bars[RU_BAR_RDY + i], RO_WARPS);
        mbar_init(&bars[RU_BAR_CHUNK], 1);
        mbar_init(&bars[RU_BAR_PAR], 1);
        mbar_init(&bars[RU_BAR_PAR + 1], 1);
        fence_mbar_init();
    }
    if (warp == RO_WARPS + 1) tmem_alloc(tmem_ptr, 512);
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem = *tmem_ptr;
    const bool fine = p.prof != nullptr && p.prof_cap >= 5120 && blockIdx.x == 0;

    // register reallocation between the roles: each setmaxnreg sits inside its role's branch so that the
    // allocator sees separate budgets (at a merge point it would take the smaller one for both)
    if (warp >= RO_WARPS) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" :: "n"(RU_REGS_SERVICE));
        if (warp == RO_WARPS) {
            if (lane == 0) {
                BProducer P{ring, 0u, 0u, l2_policy_evict_last(), p.dbg, FineProf{fine ? p.prof + 2048 : nullptr, 0, 1024, false}};
                run_rollout_b<DMODEL, DH, NKB>(P, p, h, xb, yb, par, bars, tid, warp, lane);
            }
        } else if (warp == RO_WARPS + 1) {
            BMma M{ring, 0u, 0u, tmem, bars, 0u, lane, p.dbg, FineProf{(fine && lane == 0) ? p.prof + 3072 : nullptr, 0, 1024, false}};
            run_rollout_b<DMODEL, DH, NKB>(M, p, h, xb, yb, par, bars, tid, warp, lane);
        }
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" :: "n"(RU_REGS_COMPUTE));
        BCompute C{tmem, bars, 0u, 0u, warp, lane, p.dbg, FineProf{(fine && tid == 0) ? p.prof + 4096 : nullptr, 0, 1024, false}};
        run_rollout_b<DMODEL, DH, NKB>(C, p, h, xb, yb, par, bars, tid, warp, lane);
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == RO_WARPS + 1) tmem_dealloc(tmem, 512);
}

// ----------------------------------------------------------------------------
// host side
// ----------------------------------------------------------------------------
int ro_umma_plan(ROParams* p, int smem_limit, size_t* smem_bytes) {
    const int d = p->d, Ds = p->Ds, F = p->F;
    if (!(d == 128 || d == 256)) return -1;
    if (p->heads < 1 || d % p->heads) return -1;
    const int dh = d / p->heads;
    if (!((d == 128 && dh == 16) || (d == 256 && dh == 32))) return -1;
    if (Ds % 64 || F % 128 || Ds > 256 || Ds < 64) return -1;
    if (p->lmax < 1 || p->lmax > 64 || p->K > 16) return -1;
    const int Lp = (p->lmax + 15) & ~15;
    const int wa = (d > Ds ? d : Ds);
    int wy = 3 * d;
    const int nchunk = (F + 767) / 768;
    int fc = ((F + nchunk - 1) / nchunk + 127) / 128 * 128;     // hidden features per FFN chunk
    if (fc > wy) wy = fc;
    if ((3 * d / 128) * Lp > 512 || (fc / 128 + d / 128) * Lp > 512) return -1;   // TMEM columns
    if (fc / 128 > 8 || 3 * d / 128 > 8) return -1;                                  // accumulator barriers
    p->par_floats = 9 * d + F;
    const size_t x_bytes = (size_t)(wa / 64) * Lp * 128;
    const size_t y_bytes = (size_t)(wy / 64) * Lp * 128;
    const size_t h_bytes = (size_t)Lp * (d + RU_HPAD) * 4;
    const size_t par_bytes = (size_t)2 * p->par_floats * 4;
    size_t off = 0;
    p->off_a = (uint32_t)off; off += x_bytes;
    p->off_b = (uint32_t)off; off += y_bytes;
    p->off_h = (uint32_t)off; off += h_bytes;
    p->off_par = (uint32_t)off; off += par_bytes;
    p->off_bars = (uint32_t)off; off += RU_BAR_WORDS * 8;
    off = (off + 1023) / 1024 * 1024;
    // 32 KB stages (two k-adjacent tiles) halve the per-stage handshakes; 16 KB stages when room is short
    int stage_tiles = 2;
    if ((size_t)smem_limit < off + 3 * (size_t)stage_tiles * RU_TILE_BYTES) stage_tiles = 1;
    const size_t stage_bytes = (size_t)stage_tiles * RU_TILE_BYTES;
    if ((size_t)smem_limit < off + 3 * stage_bytes) return -1;
    int nstage = (int)(((size_t)smem_limit - off) / stage_bytes);
    if (nstage > 8) nstage = 8;
    p->stage_tiles = stage_tiles;
    p->off_ring = (uint32_t)off;
    p->nstage = nstage; p->par_double = 1; p->hg = p->heads; p->fc = fc;
    p->lda = 0; p->ldb = 0;
    *smem_bytes = off + (size_t)nstage * stage_bytes;
    return 0;
}

template <int DMODEL, int DH, int NKB>
static cudaError_t ro_umma_launch_t(const ROParams& p, size_t smem_bytes, cudaStream_t st) {
    auto kern = ro_umma_forward_kernel<DMODEL, DH, NKB>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
    if (e != cudaSuccess) return e;
    kern<<<p.B, RU_THREADS, smem_bytes, st>>>(p);
    return cudaGetLastError();
}

cudaError_t ro_umma_launch(const ROParams& p, size_t smem_bytes, cudaStream_t st) {
    const int Lp = (p.lmax + 15) & ~15;
    if (p.d == 128) {
        if (Lp <= 48) return ro_umma_launch_t<128, 16, 6>(p, smem_bytes, st);
        return ro_umma_launch_t<128, 16, 8>(p, smem_bytes, st);
    }
    if (Lp <= 48) return ro_umma_launch_t<256, 32, 6>(p, smem_bytes, st);
    return ro_umma_launch_t<256, 32, 8>(p, smem_bytes, st);
}

}  // namespace sfb
