// Autoregressive slot-Transformer rollout for sm_100a: ONE launch for the whole pred_len loop.
//
// Replaces reference SlotRollouter.forward (video_prediction/models/slotformer.py:85-126) and
// SingleStepSlotRollouter.forward (video_prediction/models/single_step_slotformer.py:49-90);
// the encoder layer is torch.nn.TransformerEncoderLayer(norm_first=True, activation=relu) in
// eval mode as built at slotformer.py:72-80.
//
// Engine ("resident clip, streamed weights"): each clip is independent (attention never
// crosses clips), so one CTA owns one clip for all pred_len steps.  The residual stream h
// (fp32) and the fp16 operand tiles stay in shared memory for the whole rollout; the sliding
// window is index arithmetic over [hist ; pred_out] in global memory (no torch.cat).
// Weights are fp16 copies pre-arranged as 64x64 panels in the canonical 128B-swizzled K-major
// layout (the layout both ldmatrix and tcgen05 descriptors consume); a producer warp streams
// them with 1-D TMA bulk copies through an mbarrier ring and runs ahead across GEMM and step
// boundaries (the panel order is static), so no L2 latency is exposed to the math warps.
// v0 (weights read straight from L2 by every warp) spent >70% of its time on that latency:
// profiles/r1_v0_timeline.txt.  All contractions run on tensor cores (fp32 accumulate);
// LayerNorm / softmax / bias / ReLU / positional encoding are fused into prologues/epilogues.
#include "ro_attn.cuh"
#include "ro_kernel.h"

namespace sfb {

static constexpr int RO_PANEL_BYTES = 64 * 64 * 2;      // one 64(n) x 64(k) fp16 panel
static constexpr int RO_PANEL_HALVES = 64 * 64;

// ----------------------------------------------------------------------------
// GEMM over a sub-block of a packed weight matrix
// ----------------------------------------------------------------------------
struct GemmOp {
    const __half* base;   // packed matrix
    int kpt;              // k-panels per n-panel row of the whole matrix (Kd_total / 64)
    int nb0, nnb;         // n-panel range (64 output columns each)
    int kb0, nkb;         // k-panel range
};

struct Ring {
    unsigned char* stages;
    uint64_t* full;
    uint64_t* empty;
    int nstage;
};

// Producer role: one thread, walks the same GEMM sequence as the math warps and feeds panels.
struct Producer {
    static constexpr bool kConsumer = false;
    Ring ring;
    uint32_t pidx;
    uint64_t pol;
    __device__ __forceinline__ void emit(const GemmOp& op, int nb, int kb) {
        const int s = pidx % ring.nstage;
        mbar_wait_sleep(&ring.empty[s], ((pidx / ring.nstage) & 1) ^ 1);
        mbar_arrive_expect_tx(&ring.full[s], RO_PANEL_BYTES);
        bulk_g2s(ring.stages + (size_t)s * RO_PANEL_BYTES,
                 op.base + (((size_t)((op.nb0 + nb) >> 1) * op.kpt + op.kb0 + kb) * 2 + ((op.nb0 + nb) & 1)) * RO_PANEL_HALVES,
                 RO_PANEL_BYTES, &ring.full[s], pol);
        ++pidx;
    }
    template <class Pre, class Epi>
    __device__ __forceinline__ void gemm(const GemmOp& op, const __half*, int, int, Pre, Epi) {
        for (int nbq = 0; nbq < op.nnb; nbq += 2)
            for (int kb = 0; kb < op.nkb; ++kb) {
                emit(op, nbq, kb);
                if (nbq + 1 < op.nnb) emit(op, nbq + 1, kb);
            }
    }
    __device__ __forceinline__ void sync() {}
};

// Consumer role: 8 math warps in two groups of 4; group g owns the n-panels nbq+g, warp wq of a
// group owns 16 of the panel's 64 output columns for all row blocks.
template <int NMB>
struct Consumer {
    static constexpr bool kConsumer = true;
    Ring ring;
    uint32_t pidx;
    int warp, lane;
    __device__ __forceinline__ void sync() { named_bar_sync(1, RO_THREADS); }

    // out(m, n) = sum_k A[m][k] W[n][k];  pre(col) -> float2 bias-like values loaded BEFORE the
    // k loop (latency hidden);  epi(row, col, v0, v1, pre_value) for columns (col, col+1).
    template <class Pre, class Epi>
    __device__ __forceinline__ void gemm(const GemmOp& op, const __half* A, int lda, int nmb, Pre pre, Epi epi) {
        const int grp = warp >> 2, wq = warp & 3;
        const int g = lane >> 2, t4 = lane & 3;
        const uint32_t a_u32 = smem_u32(A);
        uint32_t base = pidx;
        for (int nbq = 0; nbq < op.nnb; nbq += 2) {
            const bool pair = nbq + 1 < op.nnb;
            const bool active = (grp == 0) || pair;
            if (active) {
                const int col0 = (nbq + grp) * 64 + 16 * wq + 2 * t4;
                const float2 pv0 = pre(col0), pv1 = pre(col0 + 8);
                float acc[NMB][2][4];
#pragma unroll
                for (int mb = 0; mb < NMB; ++mb)
#pragma unroll
                    for (int nb = 0; nb < 2; ++nb) acc[mb][nb][0] = acc[mb][nb][1] = acc[mb][nb][2] = acc[mb][nb][3] = 0.f;
#pragma unroll 1
                for (int kb = 0; kb < op.nkb; ++kb) {
                    const uint32_t seq = base + (pair ? 2 * kb + grp : kb);
                    const int s = seq % ring.nstage;
                    mbar_wait(&ring.full[s], (seq / ring.nstage) & 1);
                    const uint32_t pan = smem_u32(ring.stages + (size_t)s * RO_PANEL_BYTES);
                    uint32_t bf[4][4];
                    {
                        const int row = 16 * wq + (lane & 7) + (lane >> 4) * 8;
#pragma unroll
                        for (int ks = 0; ks < 4; ++ks)
                            ldsm_x4(bf[ks], pan + row * 128 + (((2 * ks + ((lane >> 3) & 1)) ^ (row & 7)) << 4));
                    }
#pragma unroll
                    for (int mb = 0; mb < NMB; ++mb) {
                        if (mb < nmb) {
                            uint32_t af[4][4];
                            const int row = 16 * mb + (lane & 7) + ((lane >> 3) & 1) * 8;
#pragma unroll
                            for (int ks = 0; ks < 4; ++ks)
                                ldsm_x4(af[ks], a_u32 + (uint32_t)(row * lda + kb * 64 + 16 * ks + (lane >> 4) * 8) * 2u);
#pragma unroll
                            for (int ks = 0; ks < 4; ++ks) {
                                mma_f16(acc[mb][0], af[ks], bf[ks][0], bf[ks][1]);
                                mma_f16(acc[mb][1], af[ks], bf[ks][2], bf[ks][3]);
                            }
                        }
                    }
                    // The stage is released only after the panel fragments have been CONSUMED: an arrive issued right
                    // after the ldmatrix instructions does not wait for their data, and the producer's next TMA fill
                    // could overwrite the panel under a load still in flight (seen as rare, non-repeatable outputs
                    // with the d = 256 shapes).
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&ring.empty[s]);
                }
#pragma unroll
                for (int mb = 0; mb < NMB; ++mb) {
                    if (mb < nmb) {
                        const int row = 16 * mb + g;
                        epi(row, col0, acc[mb][0][0], acc[mb][0][1], pv0);
                        epi(row + 8, col0, acc[mb][0][2], acc[mb][0][3], pv0);
                        epi(row, col0 + 8, acc[mb][1][0], acc[mb][1][1], pv1);
                        epi(row + 8, col0 + 8, acc[mb][1][2], acc[mb][1][3], pv1);
                    }
                }
            }
            base += pair ? 2 * op.nkb : op.nkb;
        }
        pidx = base;
    }
};

// ----------------------------------------------------------------------------
// the rollout, written once for both roles so that their panel sequences cannot diverge
// ----------------------------------------------------------------------------
template <int DMODEL, int DH, int NKB, class Role>
__device__ __forceinline__ void run_rollout(Role& R, const ROParams& p, float* h, __half* abuf, __half* bbuf,
                                            float* par, int tid, int warp, int lane) {
    const int K = p.K, Ds = p.Ds, F = p.F;
    const int lda = p.lda, ldb = p.ldb;
    const int HG = p.hg, FC = p.fc;
    const int hgw = HG * DH;
    const float sm_scale_log2 = rsqrtf((float)DH) * 1.4426950408889634f;
    auto no_pre = [](int) { return make_float2(0.f, 0.f); };
    auto addr_a = [=](int r, int c) { return (uint32_t)((r * lda + c) * 2); };
    auto addr_b = [=](int r, int c) { return (uint32_t)((r * ldb + c) * 2); };
    const int PF = p.par_floats;
    // per-layer parameter block in smem: bqkv[3d] bo[d] b1[F] b2[d] ln1w ln1b ln2w ln2b [d each]
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
    uint32_t lcount = 0;   // layers executed so far -> parameter buffer parity
    if (Role::kConsumer) { load_params(0, 0); cp_async_wait_all(); R.sync(); }
    int pidx_prof = 0;
    const bool do_prof = Role::kConsumer && (p.prof != nullptr) && blockIdx.x == 0 && tid == 0;
    auto stamp = [&]() { if (do_prof && pidx_prof < p.prof_cap) p.prof[pidx_prof++] = globaltimer_ns(); };

    for (int b = blockIdx.x; b < p.B; b += gridDim.x) {
        const float* hist = p.hist + (size_t)b * p.hist_tokens * Ds;
        float* pred = p.pred + (size_t)b * p.pred_len * K * Ds;

        for (int step = 0; step < p.pred_len; ++step) {
            const int total = p.hist_tokens + step * K;
            int L, base, pe0;
            if (p.mode == 0) { L = p.hist_tokens; base = step * K; pe0 = 0; }
            else { L = total < p.cond_tokens ? total : p.cond_tokens; base = total - L; pe0 = p.pe_tokens - L; }
            const int Lp = (L + 15) & ~15, nmb = Lp >> 4, nkb = Lp >> 3;

            stamp();   // step start
            if (Role::kConsumer) {
                // window tokens -> fp16 A tile (rows >= L zeroed)
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
                    *reinterpret_cast<uint2*>(abuf + r * lda + c4) = pk;
                }
                R.sync();
            }
            // ---- in_proj + positional encoding -> h (slotformer.py:115-117) ----
            {
                const GemmOp op{p.w_in, Ds >> 6, 0, DMODEL >> 6, 0, Ds >> 6};
                R.gemm(op, abuf, lda, nmb,
                       [&](int col) { return __ldg(reinterpret_cast<const float2*>(p.b_in + col)); },
                       [&](int row, int col, float v0, float v1, float2 bi) {
                           float2 o = make_float2(0.f, 0.f);
                           if (row < L) {
                               const float2 pe = __ldg(reinterpret_cast<const float2*>(p.pe + (size_t)(pe0 + row) * DMODEL + col));
                               o.x = v0 + bi.x + pe.x; o.y = v1 + bi.y + pe.y;
                           }
                           *reinterpret_cast<float2*>(h + row * DMODEL + col) = o;
                       });
            }
            R.sync();
            stamp();   // in_proj done

            for (int layer = 0; layer < p.layers; ++layer) {
                const ROLayer& ly = p.layer[layer];
                const float* pb = par + (size_t)(p.par_double ? (lcount & 1) : 0) * PF;
                if (Role::kConsumer && !p.par_double && lcount > 0) {   // single buffer: load now
                    load_params(layer, 0); cp_async_wait_all(); R.sync();
                }
                const float* s_bqkv = pb;
                const float* s_bo = pb + 3 * DMODEL;
                const float* s_b1 = pb + 4 * DMODEL;
                const float* s_b2 = pb + 4 * DMODEL + F;
                const float* l1w = pb + 5 * DMODEL + F;
                const float* l1b = l1w + DMODEL;
                const float* l2w = l1w + 2 * DMODEL;
                const float* l2b = l1w + 3 * DMODEL;
                if (Role::kConsumer && p.par_double) load_params((layer + 1) % p.layers, (lcount + 1) & 1);   // prefetch next layer
                if (Role::kConsumer) {
                    ln_to_half<DMODEL, (NKB < 6 ? NKB : 6)>(h, reinterpret_cast<unsigned char*>(abuf), addr_a, L, Lp, l1w, l1b, warp, lane);
                    R.sync();
                }
                stamp();   // LN1 done
                // ---- self-attention, HG heads at a time ----
                for (int h0 = 0; h0 < p.heads; h0 += HG) {
                    if (HG == p.heads) {
                        const GemmOp op{ly.wqkv, DMODEL >> 6, 0, (3 * DMODEL) >> 6, 0, DMODEL >> 6};
                        R.gemm(op, abuf, lda, nmb,
                               [&](int col) { return *reinterpret_cast<const float2*>(s_bqkv + col); },
                               [&](int row, int col, float v0, float v1, float2 bi) {
                                   *reinterpret_cast<__half2*>(bbuf + row * ldb + col) = __floats2half2_rn(v0 + bi.x, v1 + bi.y);
                               });
                    } else {
                        for (int part = 0; part < 3; ++part) {
                            const int wrow = part * DMODEL + h0 * DH;
                            const GemmOp op{ly.wqkv, DMODEL >> 6, wrow >> 6, hgw >> 6, 0, DMODEL >> 6};
                            R.gemm(op, abuf, lda, nmb,
                                   [&](int col) { return *reinterpret_cast<const float2*>(s_bqkv + wrow + col); },
                                   [&](int row, int col, float v0, float v1, float2 bi) {
                                       *reinterpret_cast<__half2*>(bbuf + row * ldb + part * hgw + col) =
                                           __floats2half2_rn(v0 + bi.x, v1 + bi.y);
                                   });
                        }
                    }
                    if (Role::kConsumer) {
                        R.sync();
                        stamp();   // qkv done
                        if (NKB <= 6) {
                            for (int hh = warp; hh < HG; hh += RO_WARPS)
                                attn_head<DH, (NKB <= 6 ? NKB : 2), (NKB <= 6 ? NKB / 2 : 1)>(
                                    reinterpret_cast<unsigned char*>(bbuf), addr_b, nmb, hh * DH, hgw + hh * DH, 2 * hgw + hh * DH, L, nkb, sm_scale_log2, lane);
                        } else {
                            for (int item = warp; item < HG * nmb; item += RO_WARPS) {
                                const int hh = item / nmb, mb = item % nmb;
                                attn_block<DH, NKB>(reinterpret_cast<unsigned char*>(bbuf), addr_b, mb, hh * DH, hgw + hh * DH, 2 * hgw + hh * DH, L, nkb,
                                                    sm_scale_log2, lane);
                            }
                        }
                        R.sync();
                        stamp();   // attention done
                    }
                    // h += O_group Wo[:, group]^T (+ bias once)
                    const bool first = (h0 == 0);
                    const GemmOp op{ly.wo, DMODEL >> 6, 0, DMODEL >> 6, (h0 * DH) >> 6, hgw >> 6};
                    R.gemm(op, bbuf, ldb, nmb,
                           [&](int col) { return first ? *reinterpret_cast<const float2*>(s_bo + col) : make_float2(0.f, 0.f); },
                           [&](int row, int col, float v0, float v1, float2 bi) {
                               float2* hp = reinterpret_cast<float2*>(h + row * DMODEL + col);
                               float2 cur = *hp;
                               cur.x += v0 + bi.x; cur.y += v1 + bi.y;
                               *hp = cur;
                           });
                    R.sync();
                    stamp();   // out-proj done
                }
                // ---- y = LN2(h);  h += W2 relu(W1 y + b1) + b2, FC hidden columns at a time ----
                if (Role::kConsumer) {
                    ln_to_half<DMODEL, (NKB < 6 ? NKB : 6)>(h, reinterpret_cast<unsigned char*>(abuf), addr_a, L, Lp, l2w, l2b, warp, lane);
                    R.sync();
                }
                stamp();   // LN2 done
                for (int f0 = 0; f0 < F; f0 += FC) {
                    const int fcw = (F - f0) < FC ? (F - f0) : FC;
                    {
                        const GemmOp op{ly.w1, DMODEL >> 6, f0 >> 6, fcw >> 6, 0, DMODEL >> 6};
                        R.gemm(op, abuf, lda, nmb,
                               [&](int col) { return *reinterpret_cast<const float2*>(s_b1 + f0 + col); },
                               [&](int row, int col, float v0, float v1, float2 bi) {
                                   *reinterpret_cast<__half2*>(bbuf + row * ldb + col) =
                                       __floats2half2_rn(fmaxf(v0 + bi.x, 0.f), fmaxf(v1 + bi.y, 0.f));
                               });
                    }
                    R.sync();
                    stamp();   // ffn1 chunk done
                    const bool first = (f0 == 0);
                    {
                        const GemmOp op{ly.w2, F >> 6, 0, DMODEL >> 6, f0 >> 6, fcw >> 6};
                        R.gemm(op, bbuf, ldb, nmb,
                               [&](int col) { return first ? *reinterpret_cast<const float2*>(s_b2 + col) : make_float2(0.f, 0.f); },
                               [&](int row, int col, float v0, float v1, float2 bi) {
                                   float2* hp = reinterpret_cast<float2*>(h + row * DMODEL + col);
                                   float2 cur = *hp;
                                   cur.x += v0 + bi.x; cur.y += v1 + bi.y;
                                   *hp = cur;
                               });
                    }
                    if (Role::kConsumer && p.par_double && f0 + FC >= F) cp_async_wait_all();   // next layer's parameters landed
                    R.sync();
                    stamp();   // ffn2 chunk done
                }
                ++lcount;
            }

            // ---- out_proj on the last K tokens -> pred_out[b, step] (slotformer.py:121) ----
            if (Role::kConsumer) {
                for (int i = tid; i < 16 * DMODEL; i += RO_THREADS) {
                    const int r = i / DMODEL, c = i % DMODEL;
                    abuf[r * lda + c] = __float2half_rn(r < K ? h[(L - K + r) * DMODEL + c] : 0.f);
                }
                R.sync();
            }
            float* dst = pred + (size_t)step * K * Ds;
            {
                const GemmOp op{p.w_out, DMODEL >> 6, 0, Ds >> 6, 0, DMODEL >> 6};
                R.gemm(op, abuf, lda, 1,
                       [&](int col) { return __ldg(reinterpret_cast<const float2*>(p.b_out + col)); },
                       [&](int row, int col, float v0, float v1, float2 bi) {
                           if (row < K)
                               *reinterpret_cast<float2*>(dst + (size_t)row * Ds + col) = make_float2(v0 + bi.x, v1 + bi.y);
                       });
            }
            R.sync();   // pred_out[step] is visible to this CTA's next window load
        }
    }
    (void)no_pre;
}

template <int DMODEL, int DH, int NKB>
__global__ void __launch_bounds__(RO_THREADS + 32, 1) ro_mma_forward_kernel(const ROParams p) {
    extern __shared__ __align__(1024) unsigned char smem[];
    float* h = reinterpret_cast<float*>(smem + p.off_h);        // [Lmax_p][DMODEL]
    __half* abuf = reinterpret_cast<__half*>(smem + p.off_a);   // [Lmax_p][lda]
    __half* bbuf = reinterpret_cast<__half*>(smem + p.off_b);   // [Lmax_p][ldb]
    float* par = reinterpret_cast<float*>(smem + p.off_par);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + p.off_bars);
    Ring ring{smem + p.off_ring, bars, bars + 8, p.nstage};
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (tid == 0) {
        for (int s = 0; s < p.nstage; ++s) { mbar_init(&ring.full[s], 1); mbar_init(&ring.empty[s], 4); }
        fence_mbar_init();
    }
    __syncthreads();

    if (warp == RO_WARPS) {
        if (lane == 0) {
            Producer P{ring, 0u, l2_policy_evict_last()};
            run_rollout<DMODEL, DH, NKB>(P, p, h, abuf, bbuf, par, tid, warp, lane);
        }
        return;
    }
    Consumer<NKB / 2> C{ring, 0u, warp, lane};
    run_rollout<DMODEL, DH, NKB>(C, p, h, abuf, bbuf, par, tid, warp, lane);
}

// ----------------------------------------------------------------------------
// host side
// ----------------------------------------------------------------------------
int ro_mma_plan(ROParams* p, int smem_limit, size_t* smem_bytes) {
    const int d = p->d, Ds = p->Ds, F = p->F;
    if (!(d == 128 || d == 256)) return -1;
    if (p->heads < 1 || d % p->heads) return -1;
    const int dh = d / p->heads;
    if (!((d == 128 && dh == 16) || (d == 256 && dh == 32))) return -1;
    if (Ds % 64 || F % 64 || Ds > 256 || Ds < 64) return -1;
    if (p->lmax < 1 || p->lmax > 128 || p->K > 16) return -1;
    const int Lp = (p->lmax + 15) & ~15;
    const int wa = (d > Ds ? d : Ds);
    p->lda = wa + 8;
    const size_t h_bytes = (size_t)Lp * d * 4;
    const size_t a_bytes = (size_t)Lp * p->lda * 2;
    p->par_floats = 9 * d + F;
    const size_t bar_bytes = 16 * 8;
    for (int par_double = 1; par_double >= 0; --par_double) {
        const size_t par_bytes = (size_t)(par_double ? 2 : 1) * p->par_floats * 4;
        for (int hg = p->heads; hg >= 1; hg >>= 1) {
            if (p->heads % hg) continue;
            const int hgw = hg * dh;
            if (hgw % 64) break;
            int fc = 3 * hgw;
            if (fc > F) fc = F;
            fc = fc / 64 * 64;
            const int ldb = 3 * hgw + 8;
            const size_t b_bytes = (size_t)Lp * ldb * 2;
            size_t fixed = h_bytes + a_bytes + b_bytes + bar_bytes + par_bytes;
            fixed = (fixed + 1023) / 1024 * 1024;
            if ((size_t)smem_limit < fixed + 3 * (size_t)RO_PANEL_BYTES) continue;
            int nstage = (int)(((size_t)smem_limit - fixed) / RO_PANEL_BYTES);
            if (nstage > 8) nstage = 8;
            p->hg = hg; p->fc = fc; p->ldb = ldb; p->nstage = nstage; p->par_double = par_double;
            p->off_h = 0; p->off_a = (uint32_t)h_bytes; p->off_b = (uint32_t)(h_bytes + a_bytes);
            p->off_bars = (uint32_t)(h_bytes + a_bytes + b_bytes);
            p->off_par = (uint32_t)(h_bytes + a_bytes + b_bytes + bar_bytes);
            p->off_ring = (uint32_t)fixed;
            *smem_bytes = fixed + (size_t)nstage * RO_PANEL_BYTES;
            return 0;
        }
    }
    return -1;
}

template <int DMODEL, int DH, int NKB>
static cudaError_t ro_launch_t(const ROParams& p, size_t smem_bytes, cudaStream_t st) {
    auto kern = ro_mma_forward_kernel<DMODEL, DH, NKB>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
    if (e != cudaSuccess) return e;
    kern<<<p.B, RO_THREADS + 32, smem_bytes, st>>>(p);
    return cudaGetLastError();
}

cudaError_t ro_mma_launch(const ROParams& p, size_t smem_bytes, cudaStream_t st) {
    const int Lp = (p.lmax + 15) & ~15;
    if (p.d == 128) {
        if (Lp <= 48) return ro_launch_t<128, 16, 6>(p, smem_bytes, st);
        if (Lp <= 96) return ro_launch_t<128, 16, 12>(p, smem_bytes, st);
        return ro_launch_t<128, 16, 16>(p, smem_bytes, st);
    }
    if (Lp <= 48) return ro_launch_t<256, 32, 6>(p, smem_bytes, st);
    if (Lp <= 96) return ro_launch_t<256, 32, 12>(p, smem_bytes, st);
    return ro_launch_t<256, 32, 16>(p, smem_bytes, st);
}

}  // namespace sfb
