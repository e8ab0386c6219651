// LayerNorm-to-fp16 and per-(head, 16-query block) attention used by the rollout kernel.
#pragma once
#include "common.cuh"

namespace sfb {

static constexpr int RO_WARPS = 8;            // consumer warps
static constexpr int RO_THREADS = RO_WARPS * 32;
static constexpr float RO_LN_EPS = 1e-5f;

// 2^x on the SFU, one instruction (exp2f adds denormal range handling that softmax does not need); 2^-inf = 0
__device__ __forceinline__ float fast_exp2(float x) {
    float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y;
}

// LayerNorm rows [0, L) of h (fp32, stride d) -> fp16 rows of `out` (stride ldo); rows [L, Lp) = 0
template <int DMODEL, int RPW, class Addr>
__device__ __forceinline__ void ln_to_half(const float* h, unsigned char* out, Addr addr, int L, int Lp,
                                           const float* gw, const float* gb, int warp, int lane) {
    constexpr int PER0 = DMODEL / 32;
    float gmm[PER0], bta[PER0];
#pragma unroll
    for (int i = 0; i < PER0; ++i) { gmm[i] = gw[lane + 32 * i]; bta[i] = gb[lane + 32 * i]; }
  for (int row0 = 0; row0 < Lp; row0 += RO_WARPS * RPW) {
    // warp w normalises rows w, w+8, ...; RPW = max rows per warp.  All rows are loaded first and
    // their reductions are interleaved, so the shuffle latency is paid once, not once per row.
    constexpr int PER = DMODEL / 32;
    float v[RPW][PER], s[RPW], q[RPW];
#pragma unroll
    for (int j = 0; j < RPW; ++j) {
        const int r = row0 + warp + RO_WARPS * j;
        s[j] = 0.f;
#pragma unroll
        for (int i = 0; i < PER; ++i) { v[j][i] = (r < L) ? h[r * DMODEL + lane + 32 * i] : 0.f; s[j] += v[j][i]; }
    }
#pragma unroll
    for (int o = 16; o; o >>= 1)
#pragma unroll
        for (int j = 0; j < RPW; ++j) s[j] += __shfl_xor_sync(0xffffffffu, s[j], o);
#pragma unroll
    for (int j = 0; j < RPW; ++j) {
        const float mu = s[j] * (1.f / DMODEL);
        q[j] = 0.f;
#pragma unroll
        for (int i = 0; i < PER; ++i) { v[j][i] -= mu; q[j] = fmaf(v[j][i], v[j][i], q[j]); }
    }
#pragma unroll
    for (int o = 16; o; o >>= 1)
#pragma unroll
        for (int j = 0; j < RPW; ++j) q[j] += __shfl_xor_sync(0xffffffffu, q[j], o);
#pragma unroll
    for (int j = 0; j < RPW; ++j) {
        const int r = row0 + warp + RO_WARPS * j;
        if (r < Lp) {
            const float rstd = rsqrtf(q[j] * (1.f / DMODEL) + RO_LN_EPS);
#pragma unroll
            for (int i = 0; i < PER; ++i)
                *reinterpret_cast<__half*>(out + addr(r, lane + 32 * i)) =
                    __float2half_rn(r < L ? fmaf(v[j][i] * rstd, gmm[i], bta[i]) : 0.f);
        }
    }
  }
}

// Same, one lane per 4 consecutive features: 128-bit row loads, 64-bit packed fp16 stores
// (a quarter of the load/store/address instructions of ln_to_half).
template <int DMODEL, int RPW, class Addr>
__device__ __forceinline__ void ln_to_half_v(const float* h, unsigned char* out, Addr addr, int L, int Lp,
                                             const float* gw, const float* gb, int warp, int lane) {
    constexpr int NV = DMODEL / 128;
    float4 gmm[NV], bta[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        gmm[i] = *reinterpret_cast<const float4*>(gw + 4 * lane + 128 * i);
        bta[i] = *reinterpret_cast<const float4*>(gb + 4 * lane + 128 * i);
    }
    for (int row0 = 0; row0 < Lp; row0 += RO_WARPS * RPW) {
        float4 v[RPW][NV];
        float s[RPW], q[RPW];
#pragma unroll
        for (int j = 0; j < RPW; ++j) {
            const int r = row0 + warp + RO_WARPS * j;
            s[j] = 0.f;
#pragma unroll
            for (int i = 0; i < NV; ++i) {
                v[j][i] = (r < L) ? *reinterpret_cast<const float4*>(h + r * DMODEL + 4 * lane + 128 * i)
                                  : make_float4(0.f, 0.f, 0.f, 0.f);
                s[j] += (v[j][i].x + v[j][i].y) + (v[j][i].z + v[j][i].w);
            }
        }
#pragma unroll
        for (int o = 16; o; o >>= 1)
#pragma unroll
            for (int j = 0; j < RPW; ++j) s[j] += __shfl_xor_sync(0xffffffffu, s[j], o);
#pragma unroll
        for (int j = 0; j < RPW; ++j) {
            const float mu = s[j] * (1.f / DMODEL);
            q[j] = 0.f;
#pragma unroll
            for (int i = 0; i < NV; ++i) {
                v[j][i].x -= mu; v[j][i].y -= mu; v[j][i].z -= mu; v[j][i].w -= mu;
                q[j] = fmaf(v[j][i].x, v[j][i].x, q[j]); q[j] = fmaf(v[j][i].y, v[j][i].y, q[j]);
                q[j] = fmaf(v[j][i].z, v[j][i].z, q[j]); q[j] = fmaf(v[j][i].w, v[j][i].w, q[j]);
            }
        }
#pragma unroll
        for (int o = 16; o; o >>= 1)
#pragma unroll
            for (int j = 0; j < RPW; ++j) q[j] += __shfl_xor_sync(0xffffffffu, q[j], o);
#pragma unroll
        for (int j = 0; j < RPW; ++j) {
            const int r = row0 + warp + RO_WARPS * j;
            if (r < Lp) {
                const float rstd = (r < L) ? rsqrtf(q[j] * (1.f / DMODEL) + RO_LN_EPS) : 0.f;
#pragma unroll
                for (int i = 0; i < NV; ++i) {
                    uint2 pk = make_uint2(0u, 0u);
                    if (r < L) {
                        pk.x = pack_h2(fmaf(v[j][i].x * rstd, gmm[i].x, bta[i].x), fmaf(v[j][i].y * rstd, gmm[i].y, bta[i].y));
                        pk.y = pack_h2(fmaf(v[j][i].z * rstd, gmm[i].z, bta[i].z), fmaf(v[j][i].w * rstd, gmm[i].w, bta[i].w));
                    }
                    *reinterpret_cast<uint2*>(out + addr(r, 4 * lane + 128 * i)) = pk;
                }
            }
        }
    }
}

// LayerNorm with 4 lanes per row (8 rows per warp, 64 rows per pass of the 8 warps): two shuffle steps per
// reduction and a single pass for every supported window.  Lane j of a row group owns the float4 chunks at
// features 4j + 16i; rows are HPAD floats apart beyond DMODEL so that the two rows of a quarter-warp hit
// different banks.  Four independent accumulators keep the per-lane sums off one dependent chain.
template <int DMODEL, int HPAD, class Addr>
__device__ __forceinline__ void ln_to_half_q(const float* h, unsigned char* out, Addr addr, int L, int Lp,
                                             const float* gw, const float* gb, int warp, int lane, int rbegin = 0) {
    constexpr int NV = DMODEL / 16, HS = DMODEL + HPAD;
    const int j = lane & 3, rsub = lane >> 2;
    for (int row0 = rbegin; row0 < Lp; row0 += RO_WARPS * 8) {      // rows [rbegin, Lp)
        if (row0 + warp * 8 >= Lp) break;            // warp-uniform
        const int r = row0 + warp * 8 + rsub;
        const bool valid = r < L;
        float4 v[NV];
        float s4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            v[i] = valid ? *reinterpret_cast<const float4*>(h + r * HS + 4 * j + 16 * i) : make_float4(0.f, 0.f, 0.f, 0.f);
            s4[i & 3] += (v[i].x + v[i].y) + (v[i].z + v[i].w);
        }
        float s = (s4[0] + s4[1]) + (s4[2] + s4[3]);
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        const float mu = s * (1.f / DMODEL);
        float q4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            v[i].x -= mu; v[i].y -= mu; v[i].z -= mu; v[i].w -= mu;
            q4[i & 3] = fmaf(v[i].x, v[i].x, q4[i & 3]); q4[i & 3] = fmaf(v[i].y, v[i].y, q4[i & 3]);
            q4[i & 3] = fmaf(v[i].z, v[i].z, q4[i & 3]); q4[i & 3] = fmaf(v[i].w, v[i].w, q4[i & 3]);
        }
        float q = (q4[0] + q4[1]) + (q4[2] + q4[3]);
        q += __shfl_xor_sync(0xffffffffu, q, 1);
        q += __shfl_xor_sync(0xffffffffu, q, 2);
        if (r < Lp) {
            const float rstd = valid ? rsqrtf(q * (1.f / DMODEL) + RO_LN_EPS) : 0.f;   // pad rows -> exact zeros
#pragma unroll
            for (int i = 0; i < NV; ++i) {
                const float4 g4 = *reinterpret_cast<const float4*>(gw + 4 * j + 16 * i);
                const float4 b4 = *reinterpret_cast<const float4*>(gb + 4 * j + 16 * i);
                const float z = valid ? 1.f : 0.f;
                uint2 pk;
                pk.x = pack_h2(fmaf(v[i].x * rstd, g4.x, b4.x * z), fmaf(v[i].y * rstd, g4.y, b4.y * z));
                pk.y = pack_h2(fmaf(v[i].z * rstd, g4.z, b4.z * z), fmaf(v[i].w * rstd, g4.w, b4.w * z));
                *reinterpret_cast<uint2*>(out + addr(r, 4 * j + 16 * i)) = pk;
            }
        }
    }
}

// LayerNorm with 16 lanes per row (2 rows per warp, 16 rows per pass of the 8 warps) and every pass of the window
// in flight at once: a lane owns the float4 chunks at features 4j + 64i of up to MAXP rows, whose loads, reductions
// (4 shuffle steps each, issued for all rows together) and stores are independent chains.  ln_to_half_q (4 lanes per
// row, 32 features per lane in one dependent chain) needed 2700 clocks for a 48-row window at 2 warps per
// scheduler; the work per thread is the same here, the dependent chain is a quarter as long.
template <int DMODEL, int HPAD, int MAXP, class Addr>
__device__ __forceinline__ void ln_to_half_w(const float* h, unsigned char* out, Addr addr, int L, int Lp,
                                             const float* gw, const float* gb, int warp, int lane, int rbegin = 0) {
    constexpr int NV = DMODEL / 64, HS = DMODEL + HPAD;
    const int j = lane & 15, rsub = lane >> 4;
    float4 v[MAXP][NV];
    float s[MAXP], q[MAXP];
#pragma unroll
    for (int p = 0; p < MAXP; ++p) {
        const int r = rbegin + 16 * p + 2 * warp + rsub;
        s[p] = 0.f;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            v[p][i] = (r < L) ? *reinterpret_cast<const float4*>(h + r * HS + 4 * j + 64 * i) : make_float4(0.f, 0.f, 0.f, 0.f);
            s[p] += (v[p][i].x + v[p][i].y) + (v[p][i].z + v[p][i].w);
        }
    }
#pragma unroll
    for (int o = 1; o <= 8; o <<= 1)
#pragma unroll
        for (int p = 0; p < MAXP; ++p) s[p] += __shfl_xor_sync(0xffffffffu, s[p], o);
#pragma unroll
    for (int p = 0; p < MAXP; ++p) {
        const float mu = s[p] * (1.f / DMODEL);
        q[p] = 0.f;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            v[p][i].x -= mu; v[p][i].y -= mu; v[p][i].z -= mu; v[p][i].w -= mu;
            q[p] = fmaf(v[p][i].x, v[p][i].x, q[p]); q[p] = fmaf(v[p][i].y, v[p][i].y, q[p]);
            q[p] = fmaf(v[p][i].z, v[p][i].z, q[p]); q[p] = fmaf(v[p][i].w, v[p][i].w, q[p]);
        }
    }
#pragma unroll
    for (int o = 1; o <= 8; o <<= 1)
#pragma unroll
        for (int p = 0; p < MAXP; ++p) q[p] += __shfl_xor_sync(0xffffffffu, q[p], o);
    float4 g4[NV], b4[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        g4[i] = *reinterpret_cast<const float4*>(gw + 4 * j + 64 * i);
        b4[i] = *reinterpret_cast<const float4*>(gb + 4 * j + 64 * i);
    }
#pragma unroll
    for (int p = 0; p < MAXP; ++p) {
        const int r = rbegin + 16 * p + 2 * warp + rsub;
        if (r < Lp) {
            const bool valid = r < L;
            const float rstd = valid ? rsqrtf(q[p] * (1.f / DMODEL) + RO_LN_EPS) : 0.f;   // pad rows -> exact zeros
            const float z = valid ? 1.f : 0.f;
#pragma unroll
            for (int i = 0; i < NV; ++i) {
                uint2 pk;
                pk.x = pack_h2(fmaf(v[p][i].x * rstd, g4[i].x, b4[i].x * z), fmaf(v[p][i].y * rstd, g4[i].y, b4[i].y * z));
                pk.y = pack_h2(fmaf(v[p][i].z * rstd, g4[i].z, b4[i].z * z), fmaf(v[p][i].w * rstd, g4[i].w, b4[i].w * z));
                *reinterpret_cast<uint2*>(out + addr(r, 4 * j + 64 * i)) = pk;
            }
        }
    }
}

// One (head, 16-query block) of softmax(Q K^T / sqrt(dh)) V.  Q/K/V live in `buf` (fp16, stride
// ldb) at column offsets qcol/kcol/vcol; the result overwrites the Q block it came from.
template <int DH, int NKB, class Addr>
__device__ __forceinline__ void attn_rows(unsigned char* buf, Addr addr, int qrow0, int qcol, int kcol,
                                          int vcol, int L, int nkb, float sm_scale_log2, int lane);

template <int DH, int NKB, class Addr>
__device__ __forceinline__ void attn_block(unsigned char* buf, Addr addr, int mb, int qcol, int kcol,
                                           int vcol, int L, int nkb, float sm_scale_log2, int lane) {
    attn_rows<DH, NKB>(buf, addr, 16 * mb, qcol, kcol, vcol, L, nkb, sm_scale_log2, lane);
}

// 16 query rows starting at qrow0 (any multiple of 8 that keeps the block inside the buffer)
template <int DH, int NKB, class Addr>
__device__ __forceinline__ void attn_rows(unsigned char* buf, Addr addr, int qrow0, int qcol, int kcol,
                                          int vcol, int L, int nkb, float sm_scale_log2, int lane) {
    const int g = lane >> 2, t4 = lane & 3;
    const uint32_t b_u32 = smem_u32(buf);
    uint32_t qf[DH / 16][4];
#pragma unroll
    for (int ks = 0; ks < DH / 16; ++ks) {
        const int row = qrow0 + (lane & 7) + ((lane >> 3) & 1) * 8;
        ldsm_x4(qf[ks], b_u32 + addr(row, qcol + 16 * ks + (lane >> 4) * 8));
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
                ldsm_x4(kf, b_u32 + addr(row, kcol + 16 * ks + ((lane >> 3) & 1) * 8));
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
                ldsm_x4_t(vf, b_u32 + addr(row, vcol + 8 * nb + (lane >> 4) * 8));
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
        *reinterpret_cast<__half2*>(buf + addr(qrow0 + g, col)) = __floats2half2_rn(o[nb][0] * i0, o[nb][1] * i0);
        *reinterpret_cast<__half2*>(buf + addr(qrow0 + g + 8, col)) = __floats2half2_rn(o[nb][2] * i1, o[nb][3] * i1);
    }
}

// All (<= NMB) 16-query blocks of one head in one go: K/V fragments are loaded once and the
// independent per-block chains (MMA -> shuffles -> exp2 -> MMA) interleave.  Used when the window
// is short (NMB * NKB accumulators fit in registers).
// NKBV / NMBV >= 0: the number of key blocks holding a valid key / of query blocks is known at compile time (the
// common window of a configuration): every guard below folds away and the per-block chains schedule as one block.
template <int DH, int NKB, int NMB, int NKBV = -1, int NMBV = -1, class Addr>
__device__ __forceinline__ void attn_head(unsigned char* buf, Addr addr, int nmb_rt, int qcol, int kcol, int vcol,
                                          int L, int nkb_rt, float sm_scale_log2, int lane, int qrow0 = 0) {
    const int nmb = NMBV >= 0 ? NMBV : nmb_rt;
    const int nkb = NKBV >= 0 ? NKBV : nkb_rt;
    const int g = lane >> 2, t4 = lane & 3;
    const uint32_t b_u32 = smem_u32(buf);
    uint32_t qf[NMB][DH / 16][4];
#pragma unroll
    for (int mb = 0; mb < NMB; ++mb)
#pragma unroll
        for (int ks = 0; ks < DH / 16; ++ks) {
            const int row = qrow0 + 16 * (mb < nmb ? mb : 0) + (lane & 7) + ((lane >> 3) & 1) * 8;
            ldsm_x4(qf[mb][ks], b_u32 + addr(row, qcol + 16 * ks + (lane >> 4) * 8));
        }
    float s[NMB][NKB][4];
#pragma unroll
    for (int mb = 0; mb < NMB; ++mb)
#pragma unroll
        for (int nb = 0; nb < NKB; ++nb) { s[mb][nb][0] = s[mb][nb][1] = s[mb][nb][2] = s[mb][nb][3] = 0.f; }
#pragma unroll
    for (int nb = 0; nb < NKB; nb += 2) {
        if (nb < nkb) {
#pragma unroll
            for (int ks = 0; ks < DH / 16; ++ks) {
                uint32_t kf[4];
                const int row = 8 * nb + (lane & 7) + (lane >> 4) * 8;
                ldsm_x4(kf, b_u32 + addr(row, kcol + 16 * ks + ((lane >> 3) & 1) * 8));
#pragma unroll
                for (int mb = 0; mb < NMB; ++mb) {
                    if (mb < nmb) {                 // (warp-uniform) blocks beyond nmb are not computed at all
                        mma_f16(s[mb][nb], qf[mb][ks], kf[0], kf[1]);
                        mma_f16(s[mb][nb + 1], qf[mb][ks], kf[2], kf[3]);
                    }
                }
            }
        }
    }
    float m0[NMB], m1[NMB], l0[NMB], l1[NMB];
#pragma unroll
    for (int mb = 0; mb < NMB; ++mb) {
        m0[mb] = -INFINITY; m1[mb] = -INFINITY;
#pragma unroll
        for (int nb = 0; nb < NKB; ++nb) {
            if (nb < nkb && mb < nmb) {
                if (8 * nb + 8 <= L) {        // block of valid keys only (warp-uniform): no masking work
#pragma unroll
                    for (int e = 0; e < 4; ++e) s[mb][nb][e] *= sm_scale_log2;
                } else {
                    const int c = 8 * nb + 2 * t4;
                    s[mb][nb][0] = (c < L) ? s[mb][nb][0] * sm_scale_log2 : -INFINITY;
                    s[mb][nb][1] = (c + 1 < L) ? s[mb][nb][1] * sm_scale_log2 : -INFINITY;
                    s[mb][nb][2] = (c < L) ? s[mb][nb][2] * sm_scale_log2 : -INFINITY;
                    s[mb][nb][3] = (c + 1 < L) ? s[mb][nb][3] * sm_scale_log2 : -INFINITY;
                }
                m0[mb] = fmaxf(m0[mb], fmaxf(s[mb][nb][0], s[mb][nb][1]));
                m1[mb] = fmaxf(m1[mb], fmaxf(s[mb][nb][2], s[mb][nb][3]));
            }
        }
    }
#pragma unroll
    for (int o = 1; o <= 2; o <<= 1)
#pragma unroll
        for (int mb = 0; mb < NMB; ++mb) {
            m0[mb] = fmaxf(m0[mb], __shfl_xor_sync(0xffffffffu, m0[mb], o));
            m1[mb] = fmaxf(m1[mb], __shfl_xor_sync(0xffffffffu, m1[mb], o));
        }
    uint32_t pf[NMB][NKB][2];
#pragma unroll
    for (int mb = 0; mb < NMB; ++mb) {
        l0[mb] = 0.f; l1[mb] = 0.f;
#pragma unroll
        for (int nb = 0; nb < NKB; ++nb) {
            if (nb < nkb && mb < nmb) {
                const float e0 = fast_exp2(s[mb][nb][0] - m0[mb]), e1 = fast_exp2(s[mb][nb][1] - m0[mb]);
                const float e2 = fast_exp2(s[mb][nb][2] - m1[mb]), e3 = fast_exp2(s[mb][nb][3] - m1[mb]);
                l0[mb] += e0 + e1; l1[mb] += e2 + e3;
                pf[mb][nb][0] = pack_h2(e0, e1);
                pf[mb][nb][1] = pack_h2(e2, e3);
            } else {
                pf[mb][nb][0] = pf[mb][nb][1] = 0u;
            }
        }
    }
#pragma unroll
    for (int o = 1; o <= 2; o <<= 1)
#pragma unroll
        for (int mb = 0; mb < NMB; ++mb) {
            l0[mb] += __shfl_xor_sync(0xffffffffu, l0[mb], o);
            l1[mb] += __shfl_xor_sync(0xffffffffu, l1[mb], o);
        }
    float o_[NMB][DH / 8][4];
#pragma unroll
    for (int mb = 0; mb < NMB; ++mb)
#pragma unroll
        for (int nb = 0; nb < DH / 8; ++nb) { o_[mb][nb][0] = o_[mb][nb][1] = o_[mb][nb][2] = o_[mb][nb][3] = 0.f; }
#pragma unroll
    for (int kk = 0; kk < NKB / 2; ++kk) {
        if (2 * kk < nkb) {
#pragma unroll
            for (int nb = 0; nb < DH / 8; nb += 2) {
                uint32_t vf[4];
                const int row = 16 * kk + (lane & 7) + ((lane >> 3) & 1) * 8;
                ldsm_x4_t(vf, b_u32 + addr(row, vcol + 8 * nb + (lane >> 4) * 8));
#pragma unroll
                for (int mb = 0; mb < NMB; ++mb) {
                    if (mb < nmb) {
                        const uint32_t a[4] = {pf[mb][2 * kk][0], pf[mb][2 * kk][1], pf[mb][2 * kk + 1][0], pf[mb][2 * kk + 1][1]};
                        mma_f16(o_[mb][nb], a, vf[0], vf[1]);
                        mma_f16(o_[mb][nb + 1], a, vf[2], vf[3]);
                    }
                }
            }
        }
    }
    __syncwarp();   // every lane holds its Q fragments; the Q blocks may now be overwritten
#pragma unroll
    for (int mb = 0; mb < NMB; ++mb) {
        if (mb < nmb) {
            const float i0 = 1.f / l0[mb], i1 = 1.f / l1[mb];
#pragma unroll
            for (int nb = 0; nb < DH / 8; ++nb) {
                const int col = qcol + 8 * nb + 2 * t4;
                *reinterpret_cast<__half2*>(buf + addr(qrow0 + 16 * mb + g, col)) = __floats2half2_rn(o_[mb][nb][0] * i0, o_[mb][nb][1] * i0);
                *reinterpret_cast<__half2*>(buf + addr(qrow0 + 16 * mb + g + 8, col)) = __floats2half2_rn(o_[mb][nb][2] * i1, o_[mb][nb][3] * i1);
            }
        }
    }
}

}  // namespace sfb
