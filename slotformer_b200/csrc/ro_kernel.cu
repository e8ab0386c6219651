// Autoregressive slot-Transformer rollout for sm_100a: ONE launch for the whole pred_len loop.
//
// Replaces reference SlotRollouter.forward (video_prediction/models/slotformer.py:85-126) and
// SingleStepSlotRollouter.forward (video_prediction/models/single_step_slotformer.py:49-90);
// the encoder layer is torch.nn.TransformerEncoderLayer(norm_first=True, activation=relu) in
// eval mode as built at slotformer.py:72-80.
//
// v0 engine ("resident sample"): each clip is independent (attention never crosses clips), so
// one CTA owns one clip for all pred_len steps.  The residual stream h (fp32) and the fp16
// operand tiles stay in shared memory for the whole rollout; the sliding window is index
// arithmetic over [hist ; pred_out] in global memory (no torch.cat); weights are fp16 copies
// streamed from L2.  All contractions run on tensor cores (mma.sync m16n8k16, fp32 accumulate),
// LayerNorm / softmax / bias / ReLU / positional encoding are fused into GEMM prologues and
// epilogues.  No host round trip between steps.
#include "common.cuh"
#include "ro_kernel.h"

namespace sfb {

static constexpr int RO_WARPS = 8;
static constexpr int RO_THREADS = RO_WARPS * 32;
static constexpr float RO_LN_EPS = 1e-5f;

__global__ void ro_convert_kernel(const float* __restrict__ src, __half* __restrict__ dst, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) dst[i] = __float2half_rn(src[i]);
}

cudaError_t ro_convert_launch(const float* src, __half* dst, size_t n, cudaStream_t st) {
    int blocks = (int)((n + 255) / 256);
    if (blocks > 1184) blocks = 1184;
    if (blocks < 1) blocks = 1;
    ro_convert_kernel<<<blocks, 256, 0, st>>>(src, dst, n);
    return cudaGetLastError();
}

// ----------------------------------------------------------------------------
// CTA-wide GEMM:  out(m, n) = sum_k A[m][k] * W[n][k]      (A: smem fp16, W: global fp16)
//   nmb  16-row blocks of A;  N multiple of 16;  Kd multiple of KC.
//   epi(row, col, v0, v1) receives two adjacent columns (col, col+1).
// Warp tile = (up to 4 m-blocks) x 16 columns; B fragments for KC contraction elements are
// prefetched one chunk ahead so the L2 latency overlaps the MMAs of the current chunk.
// ----------------------------------------------------------------------------
template <int KC, class Epi>
__device__ __forceinline__ void cta_gemm(const __half* A, int lda, int nmb,
                                         const __half* __restrict__ W, int ldw, int N, int Kd,
                                         Epi epi, int warp, int lane) {
    constexpr int NKS = KC / 16;
    const int g = lane >> 2, t4 = lane & 3;
    const uint32_t a_u32 = smem_u32(A);
    const int ntile = N >> 4;
    for (int tile = warp; tile < ntile; tile += RO_WARPS) {
        const int n0 = tile << 4;
        const __half* wrow0 = W + (size_t)(n0 + g) * ldw + 2 * t4;
        const __half* wrow1 = wrow0 + (size_t)8 * ldw;
        for (int mb0 = 0; mb0 < nmb; mb0 += 4) {
            const int mcnt = (nmb - mb0) < 4 ? (nmb - mb0) : 4;
            float acc[4][2][4];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 2; ++j) acc[i][j][0] = acc[i][j][1] = acc[i][j][2] = acc[i][j][3] = 0.f;
            uint32_t bcur[NKS][2][2], bnxt[NKS][2][2];
            auto load_b = [&](uint32_t (&b)[NKS][2][2], int k0) {
#pragma unroll
                for (int ks = 0; ks < NKS; ++ks) {
                    b[ks][0][0] = __ldg(reinterpret_cast<const unsigned int*>(wrow0 + k0 + 16 * ks));
                    b[ks][0][1] = __ldg(reinterpret_cast<const unsigned int*>(wrow0 + k0 + 16 * ks + 8));
                    b[ks][1][0] = __ldg(reinterpret_cast<const unsigned int*>(wrow1 + k0 + 16 * ks));
                    b[ks][1][1] = __ldg(reinterpret_cast<const unsigned int*>(wrow1 + k0 + 16 * ks + 8));
                }
            };
            auto compute = [&](const uint32_t (&b)[NKS][2][2], int k0) {
#pragma unroll
                for (int ks = 0; ks < NKS; ++ks) {
#pragma unroll
                    for (int mb = 0; mb < 4; ++mb) {
                        if (mb < mcnt) {
                            uint32_t a[4];
                            const int row = 16 * (mb0 + mb) + (lane & 7) + ((lane >> 3) & 1) * 8;
                            ldsm_x4(a, a_u32 + (uint32_t)(row * lda + k0 + 16 * ks + (lane >> 4) * 8) * 2u);
                            mma_f16(acc[mb][0], a, b[ks][0][0], b[ks][0][1]);
                            mma_f16(acc[mb][1], a, b[ks][1][0], b[ks][1][1]);
                        }
                    }
                }
            };
            load_b(bcur, 0);
            for (int k0 = 0; k0 < Kd; k0 += 2 * KC) {
                if (k0 + KC < Kd) load_b(bnxt, k0 + KC);
                compute(bcur, k0);
                if (k0 + KC < Kd) {
                    if (k0 + 2 * KC < Kd) load_b(bcur, k0 + 2 * KC);
                    compute(bnxt, k0 + KC);
                }
            }
#pragma unroll
            for (int mb = 0; mb < 4; ++mb) {
                if (mb < mcnt) {
                    const int row = 16 * (mb0 + mb) + g;
#pragma unroll
                    for (int nb = 0; nb < 2; ++nb) {
                        const int col = n0 + 8 * nb + 2 * t4;
                        epi(row, col, acc[mb][nb][0], acc[mb][nb][1]);
                        epi(row + 8, col, acc[mb][nb][2], acc[mb][nb][3]);
                    }
                }
            }
        }
    }
}

template <class Epi>
__device__ __forceinline__ void cta_gemm_any(const __half* A, int lda, int nmb,
                                             const __half* __restrict__ W, int ldw, int N, int Kd,
                                             Epi epi, int warp, int lane) {
    if ((Kd & 127) == 0) cta_gemm<128>(A, lda, nmb, W, ldw, N, Kd, epi, warp, lane);
    else cta_gemm<64>(A, lda, nmb, W, ldw, N, Kd, epi, warp, lane);
}

// LayerNorm rows [0, L) of h (fp32, stride d) -> fp16 rows of `out` (stride ldo); rows [L, Lp) = 0
template <int DMODEL>
__device__ __forceinline__ void ln_to_half(const float* h, __half* out, int ldo, int L, int Lp,
                                           const float* __restrict__ gw, const float* __restrict__ gb,
                                           int warp, int lane) {
    constexpr int PER = DMODEL / 32;
    float gmm[PER], bta[PER];
#pragma unroll
    for (int i = 0; i < PER; ++i) { gmm[i] = __ldg(gw + lane + 32 * i); bta[i] = __ldg(gb + lane + 32 * i); }
    for (int r = warp; r < Lp; r += RO_WARPS) {
        if (r < L) {
            float v[PER];
            float s = 0.f;
#pragma unroll
            for (int i = 0; i < PER; ++i) { v[i] = h[r * DMODEL + lane + 32 * i]; s += v[i]; }
#pragma unroll
            for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            const float mu = s * (1.f / DMODEL);
            float q = 0.f;
#pragma unroll
            for (int i = 0; i < PER; ++i) { v[i] -= mu; q = fmaf(v[i], v[i], q); }
#pragma unroll
            for (int o = 16; o; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
            const float rstd = rsqrtf(q * (1.f / DMODEL) + RO_LN_EPS);
#pragma unroll
            for (int i = 0; i < PER; ++i)
                out[r * ldo + lane + 32 * i] = __float2half_rn(fmaf(v[i] * rstd, gmm[i], bta[i]));
        } else {
#pragma unroll
            for (int i = 0; i < PER; ++i) out[r * ldo + lane + 32 * i] = __float2half_rn(0.f);
        }
    }
}

// One (head, 16-query block) of softmax(Q K^T / sqrt(dh)) V.  Q/K/V live in `buf` (fp16, stride
// ldb) at column offsets qcol/kcol/vcol; the result overwrites the Q block it came from.
template <int DH, int NKB>
__device__ __forceinline__ void attn_block(__half* buf, int ldb, int mb, int qcol, int kcol,
                                           int vcol, int L, int nkb, float sm_scale_log2, int lane) {
    const int g = lane >> 2, t4 = lane & 3;
    const uint32_t b_u32 = smem_u32(buf);
    uint32_t qf[DH / 16][4];
#pragma unroll
    for (int ks = 0; ks < DH / 16; ++ks) {
        const int row = 16 * mb + (lane & 7) + ((lane >> 3) & 1) * 8;
        ldsm_x4(qf[ks], b_u32 + (uint32_t)(row * ldb + qcol + 16 * ks + (lane >> 4) * 8) * 2u);
    }
    float s[NKB][4];
#pragma unroll
    for (int nb = 0; nb < NKB; ++nb) { s[nb][0] = s[nb][1] = s[nb][2] = s[nb][3] = 0.f; }
#pragma unroll
    for (int nb = 0; nb < NKB; nb += 2) {
        if (nb < nkb) {
#pragma unroll
            for (int ks = 0; ks < DH / 16; ++ks) {
                uint32_t kf[4];
                const int row = 8 * nb + (lane & 7) + (lane >> 4) * 8;
                ldsm_x4(kf, b_u32 + (uint32_t)(row * ldb + kcol + 16 * ks + ((lane >> 3) & 1) * 8) * 2u);
                mma_f16(s[nb], qf[ks], kf[0], kf[1]);
                mma_f16(s[nb + 1], qf[ks], kf[2], kf[3]);
            }
        }
    }
    // softmax over keys (rows g and g+8); keys >= L are masked
    float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
    for (int nb = 0; nb < NKB; ++nb) {
        if (nb < nkb) {
            const int c = 8 * nb + 2 * t4;
            s[nb][0] = (c < L) ? s[nb][0] * sm_scale_log2 : -INFINITY;
            s[nb][1] = (c + 1 < L) ? s[nb][1] * sm_scale_log2 : -INFINITY;
            s[nb][2] = (c < L) ? s[nb][2] * sm_scale_log2 : -INFINITY;
            s[nb][3] = (c + 1 < L) ? s[nb][3] * sm_scale_log2 : -INFINITY;
            m0 = fmaxf(m0, fmaxf(s[nb][0], s[nb][1]));
            m1 = fmaxf(m1, fmaxf(s[nb][2], s[nb][3]));
        }
    }
    m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 1)); m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 2));
    m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 1)); m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 2));
    float l0 = 0.f, l1 = 0.f;
    uint32_t pf[NKB][2];
#pragma unroll
    for (int nb = 0; nb < NKB; ++nb) {
        if (nb < nkb) {
            const float e0 = exp2f(s[nb][0] - m0), e1 = exp2f(s[nb][1] - m0);
            const float e2 = exp2f(s[nb][2] - m1), e3 = exp2f(s[nb][3] - m1);
            const __half2 h01 = __floats2half2_rn(e0, e1), h23 = __floats2half2_rn(e2, e3);
            const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
            l0 += f01.x + f01.y; l1 += f23.x + f23.y;      // normaliser from the rounded values
            pf[nb][0] = *reinterpret_cast<const uint32_t*>(&h01);
            pf[nb][1] = *reinterpret_cast<const uint32_t*>(&h23);
        } else {
            pf[nb][0] = pf[nb][1] = 0u;
        }
    }
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    // O = P V
    float o[DH / 8][4];
#pragma unroll
    for (int nb = 0; nb < DH / 8; ++nb) { o[nb][0] = o[nb][1] = o[nb][2] = o[nb][3] = 0.f; }
#pragma unroll
    for (int kk = 0; kk < NKB / 2; ++kk) {
        if (2 * kk < nkb) {
            const uint32_t a[4] = {pf[2 * kk][0], pf[2 * kk][1], pf[2 * kk + 1][0], pf[2 * kk + 1][1]};
#pragma unroll
            for (int nb = 0; nb < DH / 8; nb += 2) {
                uint32_t vf[4];
                const int row = 16 * kk + (lane & 7) + ((lane >> 3) & 1) * 8;
                ldsm_x4_t(vf, b_u32 + (uint32_t)(row * ldb + vcol + 8 * nb + (lane >> 4) * 8) * 2u);
                mma_f16(o[nb], a, vf[0], vf[1]);
                mma_f16(o[nb + 1], a, vf[2], vf[3]);
            }
        }
    }
    const float i0 = 1.f / l0, i1 = 1.f / l1;
    __syncwarp();   // every lane has its Q fragments; the Q block may now be overwritten
#pragma unroll
    for (int nb = 0; nb < DH / 8; ++nb) {
        const int col = qcol + 8 * nb + 2 * t4;
        *reinterpret_cast<__half2*>(buf + (16 * mb + g) * ldb + col) = __floats2half2_rn(o[nb][0] * i0, o[nb][1] * i0);
        *reinterpret_cast<__half2*>(buf + (16 * mb + g + 8) * ldb + col) = __floats2half2_rn(o[nb][2] * i1, o[nb][3] * i1);
    }
}

template <int DMODEL, int DH, int NKB>
__global__ void __launch_bounds__(RO_THREADS, 1) ro_forward_kernel(const ROParams p) {
    extern __shared__ __align__(128) unsigned char smem[];
    float* h = reinterpret_cast<float*>(smem + p.off_h);        // [Lmax_p][DMODEL]
    __half* abuf = reinterpret_cast<__half*>(smem + p.off_a);   // [Lmax_p][lda]
    __half* bbuf = reinterpret_cast<__half*>(smem + p.off_b);   // [Lmax_p][ldb]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int K = p.K, Ds = p.Ds, F = p.F;
    const int lda = p.lda, ldb = p.ldb;
    const int HG = p.hg, FC = p.fc;
    const int hgw = HG * DH;                  // columns per q/k/v group block
    const float sm_scale_log2 = rsqrtf((float)DH) * 1.4426950408889634f;
    int pidx = 0;
    const bool do_prof = (p.prof != nullptr) && blockIdx.x == 0 && tid == 0;
#define RO_STAMP() do { if (do_prof && pidx < p.prof_cap) p.prof[pidx++] = globaltimer_ns(); } while (0)

    for (int b = blockIdx.x; b < p.B; b += gridDim.x) {
        const float* hist = p.hist + (size_t)b * p.hist_tokens * Ds;
        float* pred = p.pred + (size_t)b * p.pred_len * K * Ds;

        for (int step = 0; step < p.pred_len; ++step) {
            // ---- window over the virtual sequence [hist ; pred] ----
            const int total = p.hist_tokens + step * K;
            int L, base, pe0;
            if (p.mode == 0) { L = p.hist_tokens; base = step * K; pe0 = 0; }
            else { L = total < p.cond_tokens ? total : p.cond_tokens; base = total - L; pe0 = p.pe_tokens - L; }
            const int Lp = (L + 15) & ~15, nmb = Lp >> 4, nkb = Lp >> 3;

            RO_STAMP();   // step start
            // ---- window tokens -> fp16 A tile (rows >= L zeroed) ----
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
            __syncthreads();

            // ---- in_proj + positional encoding -> h (slotformer.py:115-117) ----
            cta_gemm_any(abuf, lda, nmb, p.w_in, Ds, DMODEL, Ds,
                         [&](int row, int col, float v0, float v1) {
                             float2 o = make_float2(0.f, 0.f);
                             if (row < L) {
                                 const float2 bi = __ldg(reinterpret_cast<const float2*>(p.b_in + col));
                                 const float2 pe = __ldg(reinterpret_cast<const float2*>(p.pe + (size_t)(pe0 + row) * DMODEL + col));
                                 o.x = v0 + bi.x + pe.x; o.y = v1 + bi.y + pe.y;
                             }
                             *reinterpret_cast<float2*>(h + row * DMODEL + col) = o;
                         }, warp, lane);
            __syncthreads();

            RO_STAMP();   // in_proj done
            for (int layer = 0; layer < p.layers; ++layer) {
                const ROLayer& ly = p.layer[layer];
                // ---- y = LN1(h) ----
                ln_to_half<DMODEL>(h, abuf, lda, L, Lp, ly.ln1w, ly.ln1b, warp, lane);
                __syncthreads();
                RO_STAMP();   // LN1 done
                // ---- self-attention, HG heads at a time ----
                for (int h0 = 0; h0 < p.heads; h0 += HG) {
                    // q | k | v columns of this head group -> bbuf[:, 0:3*hgw]
                    for (int part = 0; part < 3; ++part) {
                        const int wrow = part * DMODEL + h0 * DH;
                        cta_gemm_any(abuf, lda, nmb, ly.wqkv + (size_t)wrow * DMODEL, DMODEL, hgw, DMODEL,
                                     [&](int row, int col, float v0, float v1) {
                                         const float2 bi = __ldg(reinterpret_cast<const float2*>(ly.bqkv + wrow + col));
                                         *reinterpret_cast<__half2*>(bbuf + row * ldb + part * hgw + col) =
                                             __floats2half2_rn(v0 + bi.x, v1 + bi.y);
                                     }, warp, lane);
                    }
                    __syncthreads();
                    RO_STAMP();   // qkv done
                    for (int item = warp; item < HG * nmb; item += RO_WARPS) {
                        const int hh = item / nmb, mb = item % nmb;
                        attn_block<DH, NKB>(bbuf, ldb, mb, hh * DH, hgw + hh * DH, 2 * hgw + hh * DH, L, nkb,
                                            sm_scale_log2, lane);
                    }
                    __syncthreads();
                    RO_STAMP();   // attention done
                    // h += O_group Wo[:, group]^T (+ bias once)
                    const bool first = (h0 == 0);
                    cta_gemm_any(bbuf, ldb, nmb, ly.wo + h0 * DH, DMODEL, DMODEL, hgw,
                                 [&](int row, int col, float v0, float v1) {
                                     float2* hp = reinterpret_cast<float2*>(h + row * DMODEL + col);
                                     float2 cur = *hp;
                                     if (first) {
                                         const float2 bi = __ldg(reinterpret_cast<const float2*>(ly.bo + col));
                                         cur.x += bi.x; cur.y += bi.y;
                                     }
                                     cur.x += v0; cur.y += v1;
                                     *hp = cur;
                                 }, warp, lane);
                    __syncthreads();
                }
                RO_STAMP();   // out-proj done
                // ---- y = LN2(h);  h += W2 relu(W1 y + b1) + b2, FC hidden columns at a time ----
                ln_to_half<DMODEL>(h, abuf, lda, L, Lp, ly.ln2w, ly.ln2b, warp, lane);
                __syncthreads();
                RO_STAMP();   // LN2 done
                for (int f0 = 0; f0 < F; f0 += FC) {
                    const int fcw = (F - f0) < FC ? (F - f0) : FC;
                    cta_gemm_any(abuf, lda, nmb, ly.w1 + (size_t)f0 * DMODEL, DMODEL, fcw, DMODEL,
                                 [&](int row, int col, float v0, float v1) {
                                     const float2 bi = __ldg(reinterpret_cast<const float2*>(ly.b1 + f0 + col));
                                     *reinterpret_cast<__half2*>(bbuf + row * ldb + col) =
                                         __floats2half2_rn(fmaxf(v0 + bi.x, 0.f), fmaxf(v1 + bi.y, 0.f));
                                 }, warp, lane);
                    __syncthreads();
                    RO_STAMP();   // ffn1 chunk done
                    const bool first = (f0 == 0);
                    cta_gemm_any(bbuf, ldb, nmb, ly.w2 + f0, F, DMODEL, fcw,
                                 [&](int row, int col, float v0, float v1) {
                                     float2* hp = reinterpret_cast<float2*>(h + row * DMODEL + col);
                                     float2 cur = *hp;
                                     if (first) {
                                         const float2 bi = __ldg(reinterpret_cast<const float2*>(ly.b2 + col));
                                         cur.x += bi.x; cur.y += bi.y;
                                     }
                                     cur.x += v0; cur.y += v1;
                                     *hp = cur;
                                 }, warp, lane);
                    __syncthreads();
                }
            }

            RO_STAMP();   // layers done
            // ---- out_proj on the last K tokens -> pred_out[b, step] (slotformer.py:121) ----
            for (int i = tid; i < 16 * DMODEL; i += RO_THREADS) {
                const int r = i / DMODEL, c = i % DMODEL;
                abuf[r * lda + c] = __float2half_rn(r < K ? h[(L - K + r) * DMODEL + c] : 0.f);
            }
            __syncthreads();
            float* dst = pred + (size_t)step * K * Ds;
            cta_gemm_any(abuf, lda, 1, p.w_out, DMODEL, Ds, DMODEL,
                         [&](int row, int col, float v0, float v1) {
                             if (row < K) {
                                 const float2 bi = __ldg(reinterpret_cast<const float2*>(p.b_out + col));
                                 *reinterpret_cast<float2*>(dst + (size_t)row * Ds + col) =
                                     make_float2(v0 + bi.x, v1 + bi.y);
                             }
                         }, warp, lane);
            __syncthreads();   // pred_out[step] is visible to this CTA's next window load
        }
    }
}

// ----------------------------------------------------------------------------
// host side
// ----------------------------------------------------------------------------
int ro_plan(ROParams* p, int smem_limit, size_t* smem_bytes) {
    const int d = p->d, Ds = p->Ds, F = p->F;
    if (!(d == 128 || d == 256)) return -1;
    if (p->heads < 1 || d % p->heads) return -1;
    const int dh = d / p->heads;
    if (!((d == 128 && dh == 16) || (d == 256 && dh == 32))) return -1;
    if (Ds % 64 || F % 64 || Ds > 256) return -1;
    if (p->lmax < 1 || p->lmax > 128) return -1;
    const int Lp = (p->lmax + 15) & ~15;
    const int wa = (d > Ds ? d : Ds);
    p->lda = wa + 8;
    const size_t h_bytes = (size_t)Lp * d * 4;
    const size_t a_bytes = (size_t)Lp * p->lda * 2;
    for (int hg = p->heads; hg >= 1; hg >>= 1) {
        if (p->heads % hg) continue;
        const int hgw = hg * dh;
        if (hgw % 64) break;
        int fc = 3 * hgw;
        if (fc > F) fc = F;
        fc = fc / 64 * 64;
        const int wb = 3 * hgw;
        const int ldb = wb + 8;
        const size_t b_bytes = (size_t)Lp * ldb * 2;
        const size_t total = h_bytes + a_bytes + b_bytes;
        if (total <= (size_t)smem_limit) {
            p->hg = hg; p->fc = fc; p->ldb = ldb;
            p->off_h = 0; p->off_a = (uint32_t)h_bytes; p->off_b = (uint32_t)(h_bytes + a_bytes);
            *smem_bytes = total;
            return 0;
        }
    }
    return -1;
}

template <int DMODEL, int DH, int NKB>
static cudaError_t ro_launch_t(const ROParams& p, size_t smem_bytes, cudaStream_t st) {
    auto kern = ro_forward_kernel<DMODEL, DH, NKB>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
    if (e != cudaSuccess) return e;
    kern<<<p.B, RO_THREADS, smem_bytes, st>>>(p);
    return cudaGetLastError();
}

cudaError_t ro_launch(const ROParams& p, size_t smem_bytes, cudaStream_t st) {
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
