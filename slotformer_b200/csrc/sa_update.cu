// Slot Attention: batched slot update (GRU + residual MLP + next q~), weight preparation and
// workspace layout.  Reference: base_slots/models/savi.py:80 (project_q), :95-100 (GRUCell, MLP).
//
// One CTA updates 16*NMB slot rows; rows are packed densely (row R of the launch = frame R / K, slot R % K), so a
// batch of K = 6 slots costs 6, not 8, rows per frame.  The fp16 hi/lo weight panels
// (64 x 64, 128B-swizzled, prepared once per call by sa_prep_kernel) are streamed by a producer
// warp with TMA bulk copies through an mbarrier ring; 8 math warps each own 8 of a panel's 64
// output columns for all rows.  Every product uses 3-term fp16 splitting
// (A_hi W_hi + A_lo W_hi + A_hi W_lo), which keeps the update at fp32-level accuracy on tensor
// cores.  The first version of this kernel read the weights straight from L2 per warp and
// was latency bound (207 us per launch, profiles/r1_sa_v1_launches.txt).
#include "common.cuh"
#include "sa_kernel.h"

namespace sfb {

static constexpr float SA_PSCALE = 1024.f;   // must match sa_pass.cu
static constexpr float LN_EPS = 1e-5f;
static constexpr int UPD_WARPS = 8;
static constexpr int UPD_THREADS = UPD_WARPS * 32;
static constexpr int PAIR_HALVES = 2 * 64 * 64;     // hi panel + lo panel
static constexpr int PAIR_BYTES = PAIR_HALVES * 2;

// ============================================================================================
// weight preparation: fold, split into fp16 hi/lo, pack into swizzled 64x64 panel pairs
// ============================================================================================
__device__ __forceinline__ void store_packed(__half* dst, int kpt, int n, int k, float v) {
    const int nb = n >> 6, r = n & 63, kb = k >> 6, kk = k & 63;
    __half* pair = dst + ((size_t)nb * kpt + kb) * PAIR_HALVES;
    const int off = r * 64 + ((((kk >> 3) ^ (r & 7)) << 3) | (kk & 7));
    const __half h = __float2half_rn(v);
    pair[off] = h;
    pair[64 * 64 + off] = __float2half_rn(v - __half2float(h));
}

__global__ void sa_prep_kernel(const float* __restrict__ wq, const float* __restrict__ wk,
                               const float* __restrict__ wv, const float* __restrict__ w_ih,
                               const float* __restrict__ w_hh, const float* __restrict__ w1,
                               const float* __restrict__ w2, const float* __restrict__ ln_w,
                               const float* __restrict__ ln_b, __half* qk, __half* iv, __half* hh,
                               __half* p1, __half* p2, int C, int D, int DM, float qscale) {
    const int n_qk = C * D, n_iv = 3 * D * C, n_hh = 3 * D * D, n_1 = DM * D, n_2 = D * DM;
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < n_qk) {                       // W_qk[c][e] = qscale * sum_d Wq[d][e] Wk[d][c]
        const int c = idx / D, e = idx % D;
        float acc = 0.f;
        for (int d = 0; d < D; ++d) acc = fmaf(wq[d * D + e], wk[d * C + c], acc);
        store_packed(qk, D >> 6, c, e, acc * qscale * ln_w[c]);     // LayerNorm gamma folded in
    } else if ((idx -= n_qk) < n_iv) {      // W_iv[j][c] = sum_d W_ih[j][d] Wv[d][c]
        const int j = idx / C, c = idx % C;
        float acc = 0.f;
        for (int d = 0; d < D; ++d) acc = fmaf(w_ih[j * D + d], wv[d * C + c], acc);
        store_packed(iv, C >> 6, j, c, acc);
    } else if ((idx -= n_iv) < n_hh) {
        store_packed(hh, D >> 6, idx / D, idx % D, w_hh[idx]);
    } else if ((idx -= n_hh) < n_1) {
        store_packed(p1, D >> 6, idx / D, idx % D, w1[idx]);
    } else if ((idx -= n_1) < n_2) {
        store_packed(p2, DM >> 6, idx / DM, idx % DM, w2[idx]);
    }
}

// wbeta[e] = qscale * sum_d Wq[d][e] (sum_c beta_c Wk[d][c]): the per-slot logit bias beta . q~[m]
// equals LNq(S)[m] . wbeta.  One block; D <= 256.
__global__ void sa_prep_wbeta_kernel(const float* __restrict__ wq, const float* __restrict__ wk,
                                     const float* __restrict__ ln_b, float* __restrict__ wbeta, int C,
                                     int D, float qscale) {
    __shared__ float tb[256];
    const int t = threadIdx.x;
    if (t < D) {
        float a = 0.f;
        for (int c = 0; c < C; ++c) a = fmaf(ln_b[c], wk[t * C + c], a);
        tb[t] = a;
    }
    __syncthreads();
    if (t < D) {
        float a = 0.f;
        for (int d = 0; d < D; ++d) a = fmaf(wq[d * D + t], tb[d], a);
        wbeta[t] = a * qscale;
    }
}

// ============================================================================================
// update kernel
// ============================================================================================
struct URing {
    unsigned char* stages;
    uint64_t* full;
    uint64_t* empty;
    int nstage;
};

struct UProducer {
    static constexpr bool kConsumer = false;
    URing ring;
    uint32_t pidx;
    uint64_t pol;
    template <class Acc>
    __device__ __forceinline__ void gemm(const __half* base, int kpt, int nb, int nkb, const __half*,
                                         const __half*, int, Acc&) {
        for (int kb = 0; kb < nkb; ++kb, ++pidx) {
            const int s = pidx % ring.nstage;
            mbar_wait_sleep(&ring.empty[s], ((pidx / ring.nstage) & 1) ^ 1);
            mbar_arrive_expect_tx(&ring.full[s], PAIR_BYTES);
            bulk_g2s(ring.stages + (size_t)s * PAIR_BYTES, base + ((size_t)nb * kpt + kb) * PAIR_HALVES,
                     PAIR_BYTES, &ring.full[s], pol);
        }
    }
    __device__ __forceinline__ void sync() {}
};

// split-precision terms of every slot-update GEMM: 3 = A_hi*W_hi + A_lo*W_hi + A_hi*W_lo (fp32-like),
// 2 = without the activation-lo term (activations rounded to fp16 once, weights still hi + lo)
#ifndef SA_UPD_TERMS
#define SA_UPD_TERMS 3
#endif

template <int NMB>
struct UConsumer {
    static constexpr bool kConsumer = true;
    URing ring;
    uint32_t pidx;
    int warp, lane;
    __device__ __forceinline__ void sync() { named_bar_sync(1, UPD_THREADS); }
    // acc[mb][e] += A[16*NMB x 64*nkb] * W[panel nb]^T  for this warp's 8 output columns
    __device__ __forceinline__ void gemm(const __half*, int, int, int nkb, const __half* Ahi,
                                         const __half* Alo, int lda, float (&acc)[NMB][4]) {
        const uint32_t ah_u32 = smem_u32(Ahi), al_u32 = smem_u32(Alo);
        // one accumulator per split term (hi*hi, lo*hi, hi*lo): three independent MMA chains per row block instead of
        // one chain of 12 dependent MMAs per k block (legacy HMMA has a ~35-cycle dependent-issue latency)
        float t3[NMB][3][4];
#pragma unroll
        for (int mb = 0; mb < NMB; ++mb)
#pragma unroll
            for (int t = 0; t < 3; ++t) t3[mb][t][0] = t3[mb][t][1] = t3[mb][t][2] = t3[mb][t][3] = 0.f;
#pragma unroll 1
        for (int kb = 0; kb < nkb; ++kb, ++pidx) {
            const int s = pidx % ring.nstage;
            mbar_wait(&ring.full[s], (pidx / ring.nstage) & 1);
            const uint32_t pan = smem_u32(ring.stages + (size_t)s * PAIR_BYTES);
            uint32_t bh[4][2], bl[4][2];
            {
                const int row = 8 * warp + (lane & 7);
#pragma unroll
                for (int hf = 0; hf < 2; ++hf) {
                    uint32_t r4[4];
                    const uint32_t off = row * 128 + (((4 * hf + (lane >> 3)) ^ (row & 7)) << 4);
                    ldsm_x4(r4, pan + off);
                    bh[2 * hf][0] = r4[0]; bh[2 * hf][1] = r4[1]; bh[2 * hf + 1][0] = r4[2]; bh[2 * hf + 1][1] = r4[3];
                    ldsm_x4(r4, pan + 64 * 64 * 2 + off);
                    bl[2 * hf][0] = r4[0]; bl[2 * hf][1] = r4[1]; bl[2 * hf + 1][0] = r4[2]; bl[2 * hf + 1][1] = r4[3];
                }
            }
#pragma unroll
            for (int mb = 0; mb < NMB; ++mb) {
                const int row = 16 * mb + (lane & 7) + ((lane >> 3) & 1) * 8;
                uint32_t ah[4][4], al[4][4];
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {
                    const uint32_t off = (uint32_t)(row * lda + kb * 64 + 16 * ks + (lane >> 4) * 8) * 2u;
                    ldsm_x4(ah[ks], ah_u32 + off);
#if SA_UPD_TERMS >= 3
                    ldsm_x4(al[ks], al_u32 + off);
#endif
                }
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {
                    mma_f16(t3[mb][0], ah[ks], bh[ks][0], bh[ks][1]);
#if SA_UPD_TERMS >= 3
                    mma_f16(t3[mb][1], al[ks], bh[ks][0], bh[ks][1]);
#endif
                    mma_f16(t3[mb][2], ah[ks], bl[ks][0], bl[ks][1]);
                }
            }
            // release the stage only once its fragments have been consumed by the MMAs above: an arrive issued right
            // after the ldmatrix instructions would not wait for their data (see ro_kernel.cu)
            __syncwarp();
            if (lane == 0) mbar_arrive(&ring.empty[s]);
        }
#pragma unroll
        for (int mb = 0; mb < NMB; ++mb)
#pragma unroll
            for (int e = 0; e < 4; ++e) acc[mb][e] += (t3[mb][0][e] + t3[mb][1][e]) + t3[mb][2][e];
    }
};

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + __expf(-x)); }

__device__ __forceinline__ void store_split(__half* hi, __half* lo, int idx, float v) {
    const __half h = __float2half_rn(v);
    hi[idx] = h;
    lo[idx] = __float2half_rn(v - __half2float(h));
}

// LayerNorm rows of `src` (fp32, stride lds) -> fp16 hi/lo operand buffers; one warp per row
template <int D>
__device__ __forceinline__ void ln_rows_split(const float* src, int lds, __half* hi, __half* lo, int lda,
                                              int rows, const float* __restrict__ gw,
                                              const float* __restrict__ gb, int warp, int lane) {
    constexpr int PER = D / 32;
    float gm[PER], bt[PER];
#pragma unroll
    for (int i = 0; i < PER; ++i) { gm[i] = gw[lane + 32 * i]; bt[i] = gb[lane + 32 * i]; }
    for (int r = warp; r < rows; r += UPD_WARPS) {
        float v[PER];
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < PER; ++i) { v[i] = src[r * lds + lane + 32 * i]; s += v[i]; }
#pragma unroll
        for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        const float mu = s * (1.f / D);
        float q = 0.f;
#pragma unroll
        for (int i = 0; i < PER; ++i) { v[i] -= mu; q = fmaf(v[i], v[i], q); }
#pragma unroll
        for (int o = 16; o; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
        const float rstd = rsqrtf(q * (1.f / D) + LN_EPS);
#pragma unroll
        for (int i = 0; i < PER; ++i)
            store_split(hi, lo, r * lda + lane + 32 * i, fmaf(v[i] * rstd, gm[i], bt[i]));
    }
}

// NSTCAP: ring depth limit.  8 = as deep as shared memory allows (one CTA per SM); 3 = a 48 KB ring, so that two
// CTAs of the 16-row variant fit one SM (used when the grid is capped: twice the CTA slots, one wave instead of 32-row CTAs)
template <int C, int NMB, int NSTCAP = 8>
struct UpdCfg {
    static constexpr int D = C, DM = 2 * C, ROWS = 16 * NMB;
    static constexpr int LDA = C + 8, LDH = DM + 8, LDS = D + 4;
    static constexpr int OFF_A = 0;                                  // a_hi | a_lo
    static constexpr int OFF_H = OFF_A + 2 * ROWS * LDA * 2;         // h_hi | h_lo (also previous slots)
    static constexpr int OFF_SP = OFF_H + 2 * ROWS * LDH * 2;        // fp32 state
    // staged parameters: b_ih[3D] b_hh[3D] b1[DM] b2[D] ln_m_w/b[2D] ln_q_w/b[2D] ln_in_w/b[2C] wbeta[D]
    static constexpr int PAR_FLOATS = 6 * D + DM + D + 2 * D + 2 * D + 2 * C + D;
    static constexpr int OFF_PAR = OFF_SP + ROWS * LDS * 4;
    static constexpr int OFF_BARS = OFF_PAR + PAR_FLOATS * 4;
    static constexpr int OFF_RING = (OFF_BARS + 2 * 8 * 8 + 1023) / 1024 * 1024;
    static constexpr int NST_FIT = (232448 - OFF_RING) / PAIR_BYTES;
    static constexpr int NST = NST_FIT > NSTCAP ? NSTCAP : NST_FIT;
    static_assert(NST >= 3, "update kernel: shared memory budget");
    static constexpr int SMEM = OFF_RING + NST * PAIR_BYTES;
};

template <int C, int NMB, int NSTCAP, class Role>
__device__ __forceinline__ void run_update(Role& R, const SAUpdateParams& p, unsigned char* smem, int tid,
                                           int warp, int lane) {
    using Cfg = UpdCfg<C, NMB, NSTCAP>;
    constexpr int D = Cfg::D, DM = Cfg::DM, ROWS = Cfg::ROWS, LDA = Cfg::LDA, LDH = Cfg::LDH, LDS = Cfg::LDS;
    __half* a_hi = reinterpret_cast<__half*>(smem + Cfg::OFF_A);
    __half* a_lo = a_hi + ROWS * LDA;
    __half* h_hi = reinterpret_cast<__half*>(smem + Cfg::OFF_H);
    __half* h_lo = h_hi + ROWS * LDH;
    float* sp = reinterpret_cast<float*>(smem + Cfg::OFF_SP);
    const float* par = reinterpret_cast<const float*>(smem + Cfg::OFF_PAR);
    const float* s_bih = par;
    const float* s_bhh = par + 3 * D;
    const float* s_b1 = par + 6 * D;
    const float* s_b2 = s_b1 + DM;
    const float* s_lnm_w = s_b2 + D;
    const float* s_lnm_b = s_lnm_w + D;
    const float* s_lnq_w = s_lnm_b + D;
    const float* s_lnq_b = s_lnq_w + D;
    const float* s_lnin_w = s_lnq_b + D;
    const float* s_lnin_b = s_lnin_w + C;
    const float* s_wbeta = s_lnin_b + C;
    const SAWeightsDev& w = p.w;
    const int K = p.K, N = p.N;
    const int g = lane >> 2, t4 = lane & 3;
    const int row0 = (int)blockIdx.x * ROWS;                 // first (dense) row of this CTA: row R = frame R / K, slot R % K
    const int total_rows = p.nframes * K;
    auto row_ok = [&](int r) { return row0 + r < total_rows; };
    auto frame_of = [&](int r) { return p.frame0 + (row0 + r) / K; };
    auto slot_of = [&](int r) { return (row0 + r) % K; };

    if (p.do_update) {
        if (Role::kConsumer) {
            // ---- u^ = (sum_chunks U / 1024 + eps*xsum) / (sum_chunks colsum / 1024 + N*eps) ----
            // (all partial-sum loads of a row are issued before the first use: one L2 latency per row)
            for (int r = warp; r < ROWS; r += UPD_WARPS) {
                const int f = frame_of(r), slot = slot_of(r);
                const bool ok = row_ok(r);
                float us[C / 32], xs[C / 32], sprev[C / 32], cs = 0.f;
#pragma unroll
                for (int i = 0; i < C / 32; ++i) { us[i] = 0.f; xs[i] = 0.f; sprev[i] = 0.f; }
                if (ok) {
                    const float* pbase = p.partials + (size_t)f * p.nchunk * p.pstride;
#pragma unroll
                    for (int i = 0; i < C / 32; ++i)
                        sprev[i] = __ldg(p.slots_prev + ((size_t)f * K + slot) * D + lane + 32 * i);
                    for (int ch0 = 0; ch0 < p.nchunk; ch0 += 4) {
                        float tu[4][C / 32], tx[4][C / 32], tc[4];
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const bool in = (ch0 + q) < p.nchunk;
                            const float* pq = pbase + (size_t)(in ? ch0 + q : ch0) * p.pstride;
                            tc[q] = in ? __ldg(pq + 8 * C + slot) : 0.f;
#pragma unroll
                            for (int i = 0; i < C / 32; ++i) {
                                tu[q][i] = in ? __ldg(pq + slot * C + lane + 32 * i) : 0.f;
                                tx[q][i] = (in && p.first) ? __ldg(pq + 8 * C + 8 + lane + 32 * i) : 0.f;
                            }
                        }
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            cs += tc[q];
#pragma unroll
                            for (int i = 0; i < C / 32; ++i) { us[i] += tu[q][i]; xs[i] += tx[q][i]; }
                        }
                    }
                }
                const float csn = cs * (1.f / SA_PSCALE);
                const float den = ok ? csn + (float)N * p.eps : 1.f;
#pragma unroll
                for (int i = 0; i < C / 32; ++i) {
                    const int c = lane + 32 * i;
                    float u = 0.f;
                    if (ok) {
                        float xsv = xs[i];
                        if (p.first) { if (slot == 0) p.xsum[(size_t)f * C + c] = xsv; }
                        else xsv = p.xsum[(size_t)f * C + c];
                        // t-statistics -> x^ statistics: x^ = gamma*t + beta
                        const float gmc = s_lnin_w[c], btc = s_lnin_b[c];
                        const float xs_hat = fmaf(gmc, xsv, (float)N * btc);
                        const float us_hat = fmaf(gmc, us[i] * (1.f / SA_PSCALE), btc * csn);
                        u = (us_hat + p.eps * xs_hat) / den;
                    }
                    store_split(a_hi, a_lo, r * LDA + c, u);
                    store_split(h_hi, h_lo, r * LDH + c, sprev[i]);     // previous slots as GEMM operand
                    sp[r * LDS + c] = sprev[i];
                }
            }
            R.sync();
        }
        // ---- GRU, 64 gate columns at a time: r,z accumulate W_iv u^ + W_hh s; n keeps both parts ----
        for (int jb = 0; jb < D / 64; ++jb) {
            float ar[NMB][4], az[NMB][4], ani[NMB][4], anh[NMB][4];
#pragma unroll
            for (int mb = 0; mb < NMB; ++mb)
#pragma unroll
                for (int e = 0; e < 4; ++e) { ar[mb][e] = az[mb][e] = ani[mb][e] = anh[mb][e] = 0.f; }
            R.gemm(w.w_iv, C / 64, jb, C / 64, a_hi, a_lo, LDA, ar);
            R.gemm(w.w_hh, D / 64, jb, D / 64, h_hi, h_lo, LDH, ar);
            R.gemm(w.w_iv, C / 64, D / 64 + jb, C / 64, a_hi, a_lo, LDA, az);
            R.gemm(w.w_hh, D / 64, D / 64 + jb, D / 64, h_hi, h_lo, LDH, az);
            R.gemm(w.w_iv, C / 64, 2 * (D / 64) + jb, C / 64, a_hi, a_lo, LDA, ani);
            R.gemm(w.w_hh, D / 64, 2 * (D / 64) + jb, D / 64, h_hi, h_lo, LDH, anh);
            if (Role::kConsumer) {
                const int col0 = 64 * jb + 8 * warp + 2 * t4;
                float bir[2], bhr[2], biz[2], bhz[2], bin_[2], bhn[2];
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    bir[q] = s_bih[col0 + q]; bhr[q] = s_bhh[col0 + q];
                    biz[q] = s_bih[D + col0 + q]; bhz[q] = s_bhh[D + col0 + q];
                    bin_[q] = s_bih[2 * D + col0 + q]; bhn[q] = s_bhh[2 * D + col0 + q];
                }
#pragma unroll
                for (int mb = 0; mb < NMB; ++mb)
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const int row = 16 * mb + g + 8 * (e >> 1), q = e & 1, col = col0 + q;
                        const float r_ = sigmoidf_(ar[mb][e] + bir[q] + bhr[q]);
                        const float z_ = sigmoidf_(az[mb][e] + biz[q] + bhz[q]);
                        const float n_ = tanhf(ani[mb][e] + bin_[q] + r_ * (anh[mb][e] + bhn[q]));
                        const float prev = sp[row * LDS + col];
                        sp[row * LDS + col] = (1.f - z_) * n_ + z_ * prev;     // s'
                    }
            }
        }
        if (Role::kConsumer) {
            R.sync();
            ln_rows_split<D>(sp, LDS, a_hi, a_lo, LDA, ROWS, s_lnm_w, s_lnm_b, warp, lane);
            R.sync();
        }
        // ---- MLP hidden: h = relu(LN(s') W1^T + b1) ----
        for (int nb = 0; nb < DM / 64; ++nb) {
            float acc[NMB][4];
#pragma unroll
            for (int mb = 0; mb < NMB; ++mb) acc[mb][0] = acc[mb][1] = acc[mb][2] = acc[mb][3] = 0.f;
            R.gemm(w.w1, D / 64, nb, D / 64, a_hi, a_lo, LDA, acc);
            if (Role::kConsumer) {
                const int col0 = 64 * nb + 8 * warp + 2 * t4;
                const float b0 = s_b1[col0], b1v = s_b1[col0 + 1];
#pragma unroll
                for (int mb = 0; mb < NMB; ++mb)
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const int row = 16 * mb + g + 8 * (e >> 1);
                        store_split(h_hi, h_lo, row * LDH + col0 + (e & 1),
                                    fmaxf(acc[mb][e] + ((e & 1) ? b1v : b0), 0.f));
                    }
            }
        }
        if (Role::kConsumer) R.sync();
        // ---- s_new = s' + h W2^T + b2 ----
        for (int nb = 0; nb < D / 64; ++nb) {
            float acc[NMB][4];
#pragma unroll
            for (int mb = 0; mb < NMB; ++mb) acc[mb][0] = acc[mb][1] = acc[mb][2] = acc[mb][3] = 0.f;
            R.gemm(w.w2, DM / 64, nb, DM / 64, h_hi, h_lo, LDH, acc);
            if (Role::kConsumer) {
                const int col0 = 64 * nb + 8 * warp + 2 * t4;
                const float b0 = s_b2[col0], b1v = s_b2[col0 + 1];
#pragma unroll
                for (int mb = 0; mb < NMB; ++mb)
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const int row = 16 * mb + g + 8 * (e >> 1), col = col0 + (e & 1);
                        const float sn = sp[row * LDS + col] + acc[mb][e] + ((e & 1) ? b1v : b0);
                        sp[row * LDS + col] = sn;
                        if (row_ok(row))
                            p.slots_out[((size_t)p.frame0 * K + row0 + row) * D + col] = sn;
                    }
            }
        }
        if (Role::kConsumer) R.sync();
    } else if (Role::kConsumer) {
        for (int r = warp; r < ROWS; r += UPD_WARPS) {
            const bool ok = row_ok(r);
#pragma unroll
            for (int i = 0; i < D / 32; ++i) {
                const int c = lane + 32 * i;
                sp[r * LDS + c] = ok ? __ldg(p.slots_prev + ((size_t)p.frame0 * K + row0 + r) * D + c) : 0.f;
            }
        }
        R.sync();
    }

    if (p.do_q) {
        // ---- q~ = LNq(S) W_qk^T -> fp16 hi/lo, rows >= K are zero ----
        if (Role::kConsumer) {
            ln_rows_split<D>(sp, LDS, a_hi, a_lo, LDA, ROWS, s_lnq_w, s_lnq_b, warp, lane);
            __syncwarp();
            // per-slot logit bias  beta . q~[m]  =  LNq(S)[m] . wbeta   (rows of this warp)
            for (int r = warp; r < ROWS; r += UPD_WARPS) {
                float a = 0.f;
#pragma unroll
                for (int i = 0; i < D / 32; ++i) {
                    const int e = lane + 32 * i;
                    a = fmaf(__half2float(a_hi[r * LDA + e]) + __half2float(a_lo[r * LDA + e]), s_wbeta[e], a);
                }
#pragma unroll
                for (int o = 16; o; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
                if (lane == 0 && row_ok(r)) {
                    float* lb = reinterpret_cast<float*>(p.qt + (size_t)frame_of(r) * p.qt_stride + 2 * 8 * C);
                    const int slot = slot_of(r);
                    lb[slot] = a;
                    if (slot == K - 1)                      // the frame's last slot also clears the padded slots
                        for (int z = K; z < 8; ++z) lb[z] = 0.f;
                }
            }
            R.sync();
        }
        for (int nb = 0; nb < C / 64; ++nb) {
            float acc[NMB][4];
#pragma unroll
            for (int mb = 0; mb < NMB; ++mb) acc[mb][0] = acc[mb][1] = acc[mb][2] = acc[mb][3] = 0.f;
            R.gemm(w.w_qk, D / 64, nb, D / 64, a_hi, a_lo, LDA, acc);
            if (Role::kConsumer) {
                const int col0 = 64 * nb + 8 * warp + 2 * t4;
#pragma unroll
                for (int mb = 0; mb < NMB; ++mb)
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const int row = 16 * mb + g + 8 * (e >> 1), col = col0 + (e & 1);
                        if (row_ok(row)) {
                            const float v = acc[mb][e];
                            __half* qf = p.qt + (size_t)frame_of(row) * p.qt_stride;
                            const __half hv = __float2half_rn(v);
                            const __half lv = __float2half_rn(v - __half2float(hv));
                            const int slot = slot_of(row);
                            // the frame's last slot also writes the zero rows of the padded slots K .. 7
                            const int rlast = (slot == K - 1) ? 7 : slot;
                            for (int r = slot; r <= rlast; ++r) {
                                const __half hw = (r == slot) ? hv : __float2half_rn(0.f);
                                const __half lw = (r == slot) ? lv : __float2half_rn(0.f);
                                if (p.qt_swz) {
                                    // tcgen05 B operand image (sa_pass_tc.cu): rows 0-7 = hi, 8-15 = lo, two 64-channel
                                    // K-major panels of 16 rows x 128 B, 16-byte chunks XOR-swizzled by the row
                                    const int off = (col >> 6) * 1024 + ((((col & 63) >> 3) ^ r) << 3) + (col & 7);
                                    qf[off + r * 64] = hw;
                                    qf[off + (8 + r) * 64] = lw;
                                } else {
                                    qf[r * C + col] = hw;
                                    qf[8 * C + r * C + col] = lw;
                                }
                            }
                        }
                    }
            }
        }
    }
}

template <int C, int NMB, int NSTCAP = 8>
__global__ void __launch_bounds__(UPD_THREADS + 32, NSTCAP < 8 ? 2 : 1) sa_update_kernel(const SAUpdateParams p) {
    using Cfg = UpdCfg<C, NMB, NSTCAP>;
    extern __shared__ __align__(1024) unsigned char smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::OFF_BARS);
    URing ring{smem + Cfg::OFF_RING, bars, bars + 8, Cfg::NST};
    if (tid == 0) {
        for (int s = 0; s < Cfg::NST; ++s) { mbar_init(&ring.full[s], 1); mbar_init(&ring.empty[s], UPD_WARPS); }
        fence_mbar_init();
    }
    {
        constexpr int D = Cfg::D, DM = Cfg::DM;
        float* par = reinterpret_cast<float*>(smem + Cfg::OFF_PAR);
        const float* srcs[11] = {p.w.b_ih, p.w.b_hh, p.w.b1, p.w.b2, p.w.ln_m_w, p.w.ln_m_b, p.w.ln_q_w, p.w.ln_q_b,
                                 p.w.ln_in_w, p.w.ln_in_b, p.w.wbeta};
        const int lens[11] = {3 * D, 3 * D, DM, D, D, D, D, D, C, C, D};
        int off = 0;
        for (int sgm = 0; sgm < 11; ++sgm) {
            for (int i = tid; i < lens[sgm]; i += UPD_THREADS + 32) par[off + i] = __ldg(srcs[sgm] + i);
            off += lens[sgm];
        }
    }
    __syncthreads();
    if (warp == UPD_WARPS) {
        if (lane == 0) {
            UProducer P{ring, 0u, l2_policy_evict_last()};
            run_update<C, NMB, NSTCAP>(P, p, smem, tid, warp, lane);
        }
        return;
    }
    UConsumer<NMB> Cn{ring, 0u, warp, lane};
    run_update<C, NMB, NSTCAP>(Cn, p, smem, tid, warp, lane);
}

// ============================================================================================
// host side
// ============================================================================================
static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

void sa_workspace_layout(int B, int chunk_frames, int N, int C, int D, int DM, int n_iter, SAWorkspace* ws) {
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 256); return o; };
    ws->w_qk = take((size_t)C * D * 2 * 2);
    ws->w_iv = take((size_t)3 * D * C * 2 * 2);
    ws->w_hh = take((size_t)3 * D * D * 2 * 2);
    ws->w1 = take((size_t)DM * D * 2 * 2);
    ws->w2 = take((size_t)D * DM * 2 * 2);
    ws->wbeta = take((size_t)D * 4);
    // items of 16 px x 8 warps (pass kernel) x up to 8 tiles per warp
    const int unit = 16 * 8;
    int chunk_px = 8 * unit;
    while (chunk_px > unit && chunk_px / 2 >= N) chunk_px /= 2;
    ws->chunk_px = chunk_px;
    ws->nchunk = (N + chunk_px - 1) / chunk_px;
    ws->n16 = ws->nchunk * chunk_px;
    ws->pstride = 9 * C + 8;
    ws->qt_stride = 2 * 8 * C + 16;
    ws->qt = take((size_t)B * ws->qt_stride * 2);
    ws->partials = take((size_t)B * ws->nchunk * ws->pstride * 4);
    ws->xsum = take((size_t)B * C * 4);
    ws->xhat_frames = (n_iter > 1) ? (chunk_frames < B ? chunk_frames : B) : 0;
    ws->xhat = take((size_t)ws->xhat_frames * ws->n16 * C * 2);
    ws->total = off;
}

cudaError_t sa_prep_launch(const float* wq, const float* wk, const float* wv, const float* w_ih,
                           const float* w_hh, const float* w1, const float* w2, const float* ln_in_w,
                           const float* ln_in_b, char* base,
                           const SAWorkspace& ws, int C, int D, int DM, cudaStream_t st) {
    const int total = C * D + 3 * D * C + 3 * D * D + DM * D + D * DM;
    const float qscale = (1.0f / sqrtf((float)D)) * 1.4426950408889634f;
    auto H = [&](size_t off) { return reinterpret_cast<__half*>(base + off); };
    sa_prep_kernel<<<(total + 255) / 256, 256, 0, st>>>(wq, wk, wv, w_ih, w_hh, w1, w2, ln_in_w, ln_in_b, H(ws.w_qk),
                                                       H(ws.w_iv), H(ws.w_hh), H(ws.w1), H(ws.w2),
                                                       C, D, DM, qscale);
    sa_prep_wbeta_kernel<<<1, 256, 0, st>>>(wq, wk, ln_in_b, reinterpret_cast<float*>(base + ws.wbeta), C, D, qscale);
    return cudaGetLastError();
}

template <int C, int NMB, int NSTCAP = 8>
static cudaError_t update_launch_t(const SAUpdateParams& p, cudaStream_t st) {
    using Cfg = UpdCfg<C, NMB, NSTCAP>;
    auto kern = sa_update_kernel<C, NMB, NSTCAP>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM);
    if (e != cudaSuccess) return e;
    const int grid = (p.nframes * p.K + Cfg::ROWS - 1) / Cfg::ROWS;      // dense rows: frame R / K, slot R % K
    kern<<<grid, UPD_THREADS + 32, Cfg::SMEM, st>>>(p);
    return cudaGetLastError();
}

cudaError_t sa_update_launch(const SAUpdateParams& p, int C, int sms, cudaStream_t st, bool capped) {
    // 16*NMB rows per CTA.  A CTA's time grows with NMB (about 13 us + 14 us per 16 rows), every CTA streams the
    // whole weight set, so: the smallest NMB whose grid still fits one wave of the SMs this call may use.
    const int need = (p.nframes * p.K + sms - 1) / sms;           // rows per CTA for a single wave
    if (C == 128) {
        if (need <= 16) return update_launch_t<128, 1>(p, st);
        // capped grid (the batch pipeline): two 16-row CTAs per SM (48 KB weight ring each) still make one wave where
        // 32-row CTAs would be needed otherwise (41 -> ~30 us per launch at 84 SMs)
        if (need <= 32 && capped) return update_launch_t<128, 1, 3>(p, st);
        if (need <= 32) return update_launch_t<128, 2>(p, st);
        if (need <= 48) return update_launch_t<128, 3>(p, st);
        return update_launch_t<128, 4>(p, st);
    }
    if (need <= 16) return update_launch_t<192, 1>(p, st);
    return update_launch_t<192, 2>(p, st);
}

}  // namespace sfb
