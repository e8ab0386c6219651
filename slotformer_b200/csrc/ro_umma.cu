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
static constexpr int RU_STAGE_TILES = 2;                    // k-adjacent tiles per ring stage
static constexpr int RU_STAGE_BYTES = RU_STAGE_TILES * RU_TILE_BYTES;
static constexpr int RU_TILE_HALVES = 8192;
static constexpr int RU_SYNC_THREADS = RO_THREADS + 32;   // compute warps + MMA warp
static constexpr int RU_THREADS = RO_THREADS + 64;

struct UOp {
    const __half* base;   // packed matrix (ro_pack2 layout)
    int kpt;              // 64-wide k blocks per feature tile of the whole matrix
    int tile0, ntile;     // 128-feature tiles
    int kb0, nkb;         // k-block range
};

struct BRing {
    unsigned char* stages;
    uint64_t* full;
    uint64_t* empty;
    int nstage;
};

struct BProducer {
    static constexpr bool kCompute = false;
    BRing ring;
    uint32_t pidx;
    uint64_t pol;
    int dbg;
    template <class Pre, class Epi>
    __device__ __forceinline__ void gemm(const UOp& op, uint32_t, int, int, Pre, Epi) {
        for (int t = 0; t < op.ntile; ++t)
            for (int kb = 0; kb < op.nkb; kb += RU_STAGE_TILES, ++pidx) {
                const int nk = (op.nkb - kb) < RU_STAGE_TILES ? (op.nkb - kb) : RU_STAGE_TILES;
                const int s = pidx % ring.nstage;
                mbar_wait_sleep(&ring.empty[s], ((pidx / ring.nstage) & 1) ^ 1);
                if (dbg & 1) { mbar_arrive(&ring.full[s]); continue; }
                mbar_arrive_expect_tx(&ring.full[s], nk * RU_TILE_BYTES);
                bulk_g2s(ring.stages + (size_t)s * RU_STAGE_BYTES,
                         op.base + ((size_t)(op.tile0 + t) * op.kpt + op.kb0 + kb) * RU_TILE_HALVES,
                         nk * RU_TILE_BYTES, &ring.full[s], pol);
            }
    }
    __device__ __forceinline__ void sync() {}
};

struct BMma {
    static constexpr bool kCompute = false;
    BRing ring;
    uint32_t pidx;
    uint32_t tmem;
    uint64_t* accfull;
    int lane;
    int dbg;
    // b_u32: smem address of the token operand (k-block 0), kblock_bytes apart per 64 k
    template <class Pre, class Epi>
    __device__ __forceinline__ void gemm(const UOp& op, uint32_t b_u32, int kblock_bytes, int ntok, Pre, Epi) {
        if (lane == 0) {
            const uint32_t idesc = umma_idesc_f16(128, ntok);
            for (int t = 0; t < op.ntile; ++t)
                for (int kb = 0; kb < op.nkb; kb += RU_STAGE_TILES, ++pidx) {
                    const int nk = (op.nkb - kb) < RU_STAGE_TILES ? (op.nkb - kb) : RU_STAGE_TILES;
                    const int s = pidx % ring.nstage;
                    mbar_wait_sleep(&ring.full[s], (pidx / ring.nstage) & 1);
                    tcgen05_fence_after();
                    // descriptors are built once per stage; a K step of 16 (32 B) or the next k tile only
                    // advances the 16-byte-granular start-address field
                    const uint64_t da0 = umma_smem_desc(smem_u32(ring.stages + (size_t)s * RU_STAGE_BYTES));
                    const uint64_t db0 = umma_smem_desc(b_u32 + kb * kblock_bytes);
                    const uint32_t dtm = tmem + (uint32_t)(t * ntok);
                    if (!(dbg & 2)) {
#pragma unroll
                        for (int kk = 0; kk < RU_STAGE_TILES; ++kk) {
                            if (kk < nk) {
                                const uint64_t dbk = db0 + (uint64_t)((kk * kblock_bytes) >> 4);
#pragma unroll
                                for (int k4 = 0; k4 < 4; ++k4)
                                    umma_f16(dtm, da0 + (uint64_t)((kk * RU_TILE_BYTES + k4 * 32) >> 4),
                                             dbk + (uint64_t)((k4 * 32) >> 4), idesc, (kb | kk | k4) != 0);
                            }
                        }
                    }
                    if (dbg & 8) mbar_arrive(&ring.empty[s]);   // (timing experiment only: frees the stage early)
                    else umma_commit(&ring.empty[s]);           // stage is free once these MMAs have read it
                }
            umma_commit(accfull);                     // accumulators of this GEMM are complete
        }
        __syncwarp();
    }
    __device__ __forceinline__ void sync() {
        named_bar_sync(1, RU_SYNC_THREADS);
        tcgen05_fence_after();
    }
};

struct BCompute {
    static constexpr bool kCompute = true;
    uint32_t tmem;
    uint64_t* accfull;
    uint32_t ngemm;
    int warp, lane;
    int dbg;
    // epi(feature, token8, values[8], pre(feature)): 8 consecutive tokens (token8 % 8 == 0) of one feature;
    // features of the warp's TMEM lane quadrant, one half of the tokens (warps w and w+4 share lanes)
    template <class Pre, class Epi>
    __device__ __forceinline__ void gemm(const UOp& op, uint32_t, int, int ntok, Pre pre, Epi epi) {
        mbar_wait(accfull, ngemm & 1);
        ++ngemm;
        tcgen05_fence_after();
        const int q = warp & 3, half = ntok >> 1, t0 = (warp >> 2) * half;
        for (int t = 0; t < ((dbg & 4) ? 0 : op.ntile); ++t) {
            const int f = t * 128 + 32 * q + lane;
            const float pv = pre(f);
            const uint32_t ta = tmem + ((uint32_t)(32 * q) << 16) + (uint32_t)(t * ntok + t0);
            for (int c0 = 0; c0 < half; c0 += 16) {
                float v0[8], v1[8];
                tmem_ld8(ta + c0, v0);
                const bool two = (c0 + 8) < half;
                if (two) tmem_ld8(ta + c0 + 8, v1);
                tmem_ld_wait();
                epi(f, t0 + c0, v0, pv);                 // 8 consecutive tokens, base a multiple of 8
                if (two) epi(f, t0 + c0 + 8, v1, pv);
            }
        }
        tcgen05_fence_before();
    }
    __device__ __forceinline__ void sync() {
        if (!(dbg & 16)) fence_proxy_async();   // operand tiles written by these threads -> tensor-core reads
        named_bar_sync(1, RU_SYNC_THREADS);
    }
};

template <int DMODEL, int DH, int NKB, class Role>
__device__ __forceinline__ void run_rollout_b(Role& R, const ROParams& p, float* h, unsigned char* xb,
                                              unsigned char* yb, float* par, int tid, int warp, int lane) {
    const int K = p.K, Ds = p.Ds, F = p.F, FC = p.fc;
    const int LpMax = (p.lmax + 15) & ~15;
    const int kbb = LpMax * 128;                       // bytes between 64-wide k blocks of X / Y
    const uint32_t x_u32 = smem_u32(xb), y_u32 = smem_u32(yb);
    const float sm_scale_log2 = rsqrtf((float)DH) * 1.4426950408889634f;
    auto addr_s = [=](int r, int c) { return swz_off(r, c, kbb); };
    const int PF = p.par_floats;
    auto load_params = [&](int layer, int buf) {
        const ROLayer& ly = p.layer[layer];
        float* dst = par + (size_t)buf * PF;
        const float* srcs[8] = {ly.bqkv, ly.bo, ly.b1, ly.b2, ly.ln1w, ly.ln1b, ly.ln2w, ly.ln2b};
        const int lens[8] = {3 * DMODEL, DMODEL, F, DMODEL, DMODEL, DMODEL, DMODEL, DMODEL};
        int off = 0;
        for (int sgm = 0; sgm < 8; ++sgm) {
            for (int i = tid * 4; i < lens[sgm]; i += RO_THREADS * 4) cp_async16(dst + off + i, srcs[sgm] + i);
            off += lens[sgm];
        }
        cp_async_commit();
    };
    int pidx_prof = 0;
    const bool do_prof = Role::kCompute && (p.prof != nullptr) && blockIdx.x == 0 && tid == 0;
    auto stamp = [&]() { if (do_prof && pidx_prof < p.prof_cap) p.prof[pidx_prof++] = globaltimer_ns(); };
    uint32_t lcount = 0;
    if (Role::kCompute) { load_params(0, 0); cp_async_wait_all(); }
    R.sync();

    for (int b = blockIdx.x; b < p.B; b += gridDim.x) {
        const float* hist = p.hist + (size_t)b * p.hist_tokens * Ds;
        float* pred = p.pred + (size_t)b * p.pred_len * K * Ds;
        for (int step = 0; step < p.pred_len; ++step) {
            const int total = p.hist_tokens + step * K;
            int L, base, pe0;
            if (p.mode == 0) { L = p.hist_tokens; base = step * K; pe0 = 0; }
            else { L = total < p.cond_tokens ? total : p.cond_tokens; base = total - L; pe0 = p.pe_tokens - L; }
            const int Lp = (L + 15) & ~15, nmb = Lp >> 4, nkb = Lp >> 3;

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
                    uint2 pk; pk.x = pack_h2(v.x, v.y); pk.y = pack_h2(v.z, v.w);
                    *reinterpret_cast<uint2*>(xb + addr_s(r, c4)) = pk;
                }
            }
            R.sync();
            // ---- in_proj + positional encoding -> h ----
            {
                const UOp op{p.w_in, Ds >> 6, 0, DMODEL >> 7, 0, Ds >> 6};
                R.gemm(op, x_u32, kbb, Lp,
                       [&](int f) { return __ldg(p.b_in + f); },
                       [&](int f, int t8, const float (&v)[8], float bi) {
#pragma unroll
                           for (int i = 0; i < 8; ++i) {
                               const int t = t8 + i;
                               h[t * DMODEL + f] = (t < L) ? v[i] + bi + __ldg(p.pe + (size_t)(pe0 + t) * DMODEL + f) : 0.f;
                           }
                       });
            }
            R.sync();
            stamp();   // in_proj done

            for (int layer = 0; layer < p.layers; ++layer) {
                const ROLayer& ly = p.layer[layer];
                const float* pb = par + (size_t)(p.par_double ? (lcount & 1) : 0) * PF;
                if (!p.par_double && lcount > 0) {
                    if (Role::kCompute) { load_params(layer, 0); cp_async_wait_all(); }
                    R.sync();
                }
                const float* s_bqkv = pb;
                const float* s_bo = pb + 3 * DMODEL;
                const float* s_b1 = pb + 4 * DMODEL;
                const float* s_b2 = pb + 4 * DMODEL + F;
                const float* l1w = pb + 5 * DMODEL + F;
                const float* l1b = l1w + DMODEL;
                const float* l2w = l1w + 2 * DMODEL;
                const float* l2b = l1w + 3 * DMODEL;
                if (Role::kCompute) {
                    if (p.par_double) load_params((layer + 1) % p.layers, (lcount + 1) & 1);
                    if (!(p.dbg & 64)) ln_to_half<DMODEL, (NKB < 6 ? NKB : 6)>(h, xb, addr_s, L, Lp, l1w, l1b, warp, lane);
                }
                R.sync();
                stamp();   // LN1
                // ---- packed q | k | v projection -> Y[token][3d] ----
                {
                    const UOp op{ly.wqkv, DMODEL >> 6, 0, (3 * DMODEL) >> 7, 0, DMODEL >> 6};
                    R.gemm(op, x_u32, kbb, Lp,
                           [&](int f) { return s_bqkv[f]; },
                           [&](int f, int t8, const float (&v)[8], float bi) {
                               unsigned char* yf = yb + (f >> 6) * kbb + t8 * 128 + (f & 7) * 2;
                               const int chunk = (f & 63) >> 3;
#pragma unroll
                               for (int i = 0; i < 8; ++i)
                                   *reinterpret_cast<__half*>(yf + i * 128 + ((chunk ^ i) << 4)) = __float2half_rn(v[i] + bi);
                           });
                }
                R.sync();
                stamp();   // qkv
                if (Role::kCompute && !(p.dbg & 32)) {
                    if (NKB <= 6) {
                        for (int hh = warp; hh < p.heads; hh += RO_WARPS)
                            attn_head<DH, (NKB <= 6 ? NKB : 2), (NKB <= 6 ? NKB / 2 : 1)>(
                                yb, addr_s, nmb, hh * DH, DMODEL + hh * DH, 2 * DMODEL + hh * DH, L, nkb, sm_scale_log2, lane);
                    } else {
                        for (int item = warp; item < p.heads * nmb; item += RO_WARPS) {
                            const int hh = item / nmb, mb = item % nmb;
                            attn_block<DH, NKB>(yb, addr_s, mb, hh * DH, DMODEL + hh * DH, 2 * DMODEL + hh * DH, L, nkb,
                                                sm_scale_log2, lane);
                        }
                    }
                }
                R.sync();
                stamp();   // attn
                // ---- h += O Wo^T + bo ----
                {
                    const UOp op{ly.wo, DMODEL >> 6, 0, DMODEL >> 7, 0, DMODEL >> 6};
                    R.gemm(op, y_u32, kbb, Lp,
                           [&](int f) { return s_bo[f]; },
                           [&](int f, int t8, const float (&v)[8], float bi) {
                               float* hf = h + t8 * DMODEL + f;
#pragma unroll
                               for (int i = 0; i < 8; ++i) hf[i * DMODEL] += v[i] + bi;
                           });
                }
                R.sync();
                stamp();   // outproj
                if (Role::kCompute && !(p.dbg & 64))
                    ln_to_half<DMODEL, (NKB < 6 ? NKB : 6)>(h, xb, addr_s, L, Lp, l2w, l2b, warp, lane);
                R.sync();
                stamp();   // LN2
                // ---- h += W2 relu(W1 y + b1) + b2, FC hidden features at a time ----
                for (int f0 = 0; f0 < F; f0 += FC) {
                    const int fcw = (F - f0) < FC ? (F - f0) : FC;
                    {
                        const UOp op{ly.w1, DMODEL >> 6, f0 >> 7, fcw >> 7, 0, DMODEL >> 6};
                        R.gemm(op, x_u32, kbb, Lp,
                               [&](int f) { return s_b1[f0 + f]; },
                               [&](int f, int t8, const float (&v)[8], float bi) {
                               unsigned char* yf = yb + (f >> 6) * kbb + t8 * 128 + (f & 7) * 2;
                               const int chunk = (f & 63) >> 3;
#pragma unroll
                               for (int i = 0; i < 8; ++i)
                                   *reinterpret_cast<__half*>(yf + i * 128 + ((chunk ^ i) << 4)) = __float2half_rn(fmaxf(v[i] + bi, 0.f));
                           });
                    }
                    R.sync();
                    stamp();   // ffn1
                    const bool first = (f0 == 0);
                    {
                        const UOp op{ly.w2, F >> 6, 0, DMODEL >> 7, f0 >> 6, fcw >> 6};
                        R.gemm(op, y_u32, kbb, Lp,
                               [&](int f) { return first ? s_b2[f] : 0.f; },
                               [&](int f, int t8, const float (&v)[8], float bi) {
                               float* hf = h + t8 * DMODEL + f;
#pragma unroll
                               for (int i = 0; i < 8; ++i) hf[i * DMODEL] += v[i] + bi;
                           });
                    }
                    if (Role::kCompute && p.par_double && f0 + FC >= F) cp_async_wait_all();
                    R.sync();
                    stamp();   // ffn2
                }
                ++lcount;
            }

            // ---- out_proj on the last K tokens -> pred_out[b, step] ----
            if (Role::kCompute) {
                for (int i = tid; i < 16 * DMODEL; i += RO_THREADS) {
                    const int r = i / DMODEL, c = i % DMODEL;
                    *reinterpret_cast<__half*>(xb + addr_s(r, c)) = __float2half_rn(r < K ? h[(L - K + r) * DMODEL + c] : 0.f);
                }
            }
            R.sync();
            float* dst = pred + (size_t)step * K * Ds;
            {
                const UOp op{p.w_out, DMODEL >> 6, 0, (Ds + 127) >> 7, 0, DMODEL >> 6};
                R.gemm(op, x_u32, kbb, 16,
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
    uint64_t* accfull = bars + 16;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 17);
    BRing ring{smem + p.off_ring, bars, bars + 8, p.nstage};
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (tid == 0) {
        for (int s = 0; s < p.nstage; ++s) { mbar_init(&ring.full[s], 1); mbar_init(&ring.empty[s], 1); }
        mbar_init(accfull, 1);
        fence_mbar_init();
    }
    if (warp == RO_WARPS + 1) tmem_alloc(tmem_ptr, 512);
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem = *tmem_ptr;

    if (warp == RO_WARPS) {
        if (lane == 0) {
            BProducer P{ring, 0u, l2_policy_evict_last(), p.dbg};
            run_rollout_b<DMODEL, DH, NKB>(P, p, h, xb, yb, par, tid, warp, lane);
        }
    } else if (warp == RO_WARPS + 1) {
        BMma M{ring, 0u, tmem, accfull, lane, p.dbg};
        run_rollout_b<DMODEL, DH, NKB>(M, p, h, xb, yb, par, tid, warp, lane);
    } else {
        BCompute C{tmem, accfull, 0u, warp, lane, p.dbg};
        run_rollout_b<DMODEL, DH, NKB>(C, p, h, xb, yb, par, tid, warp, lane);
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
    int fc = F < 768 ? F : 768;
    fc = fc / 128 * 128;
    if (fc > wy) wy = fc;
    if ((wy / 128) * Lp > 512) return -1;              // TMEM columns
    p->par_floats = 9 * d + F;
    const size_t x_bytes = (size_t)(wa / 64) * Lp * 128;
    const size_t y_bytes = (size_t)(wy / 64) * Lp * 128;
    const size_t h_bytes = (size_t)Lp * d * 4;
    for (int par_double = 1; par_double >= 0; --par_double) {
        const size_t par_bytes = (size_t)(par_double ? 2 : 1) * p->par_floats * 4;
        size_t off = 0;
        p->off_a = (uint32_t)off; off += x_bytes;
        p->off_b = (uint32_t)off; off += y_bytes;
        p->off_h = (uint32_t)off; off += h_bytes;
        p->off_par = (uint32_t)off; off += par_bytes;
        p->off_bars = (uint32_t)off; off += 20 * 8;
        off = (off + 1023) / 1024 * 1024;
        if ((size_t)smem_limit < off + 2 * (size_t)RU_STAGE_BYTES) continue;
        int nstage = (int)(((size_t)smem_limit - off) / RU_STAGE_BYTES);
        if (nstage > 6) nstage = 6;
        p->off_ring = (uint32_t)off;
        p->nstage = nstage; p->par_double = par_double; p->hg = p->heads; p->fc = fc;
        p->lda = 0; p->ldb = 0;
        *smem_bytes = off + (size_t)nstage * RU_STAGE_BYTES;
        return 0;
    }
    return -1;
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
