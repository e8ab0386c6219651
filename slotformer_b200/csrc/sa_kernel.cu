// Slot Attention forward for sm_100a -- one persistent thread-block CLUSTER per frame.
//
// Replaces reference SlotAttention.forward (base_slots/models/savi.py:56-102) and
// SlotAttentionWMask.forward (base_slots/models/steve.py:19-73).
//
// Design:
//   * The K/V projections are never materialised.  With x^ = LN(x):
//        logits[n,m] = scale * <x^[n] Wk^T, q[m]>      = <x^[n], q~[m]>,  q~ = LNq(S) W_qk
//        updates[m]  = sum_n w[n,m] (x^[n] Wv^T)        = (sum_n w[n,m] x^[n]) Wv^T
//     so one iteration only needs x^ (N x C) and two K x C matrices; W_qk = scale*log2e*Wq^T Wk
//     and W_iv = W_ih Wv are folded once per call by sa_fold_kernel.
//   * A cluster of CS CTAs owns one frame; CTA r owns rows_cta pixels.  Feature tiles
//     (32 px x C fp32) are streamed from HBM exactly once by a producer warp with 1-D TMA bulk
//     copies into an mbarrier ring; consumer warps LayerNorm them and keep x^ ON CHIP as an
//     fp16 slab (swizzled for ldmatrix), so iterations >= 2 never touch HBM again.
//   * Per 16 pixels a warp runs: logits (mma.sync m16n8k16, q~ split hi+lo fp16), softmax over
//     the slots with quad shuffles, movmatrix transpose of the probabilities, and the weighted
//     feature aggregation U^T[C x 8] += x^T P on tensor cores.  +eps is applied analytically:
//        sum_n (a+eps) x^ = sum_n a x^ + eps * sum_n x^ ,   sum_n (a+eps) = sum_n a + N eps.
//   * The slot update (GRU + residual MLP + next q~) is column-split over the cluster; slices
//     are all-gathered through distributed shared memory; consumer-only cluster syncs use
//     remote mbarrier arrives so the TMA producer can keep prefetching the next frame.
#include "common.cuh"
#include "sa_kernel.h"

namespace sfb {

static constexpr int SA_TILE_PX = 32;
static constexpr int SA_CONSUMER_WARPS = 8;
static constexpr int SA_CONSUMER_THREADS = SA_CONSUMER_WARPS * 32;
static constexpr int SA_THREADS = SA_CONSUMER_THREADS + 32;
static constexpr float SA_PSCALE = 1024.f;  // probabilities are stored as fp16(1024 * a)
static constexpr float LN_EPS = 1e-5f;

// ----------------------------------------------------------------------------
// weight folding: W_qk[c][e] = scale*log2e * sum_d Wq[d][e] Wk[d][c]   ([C][D])
//                 W_iv[j][c] = sum_d W_ih[j][d] Wv[d][c]               ([3D][C])
// ----------------------------------------------------------------------------
__global__ void sa_fold_kernel(const float* __restrict__ wq, const float* __restrict__ wk,
                               const float* __restrict__ wv, const float* __restrict__ w_ih,
                               float* __restrict__ w_qk, float* __restrict__ w_iv, int C, int D,
                               float qscale) {
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    int n_qk = C * D, n_iv = 3 * D * C;
    if (idx < n_qk) {
        int c = idx / D, e = idx % D;
        float acc = 0.f;
        for (int d = 0; d < D; ++d) acc = fmaf(wq[d * D + e], wk[d * C + c], acc);
        w_qk[idx] = acc * qscale;
    } else if (idx < n_qk + n_iv) {
        int i = idx - n_qk;
        int j = i / C, c = i % C;
        float acc = 0.f;
        for (int d = 0; d < D; ++d) acc = fmaf(w_ih[j * D + d], wv[d * C + c], acc);
        w_iv[i] = acc;
    }
}

// ----------------------------------------------------------------------------
// helpers
// ----------------------------------------------------------------------------
__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + __expf(-x)); }

// Consumer-only cluster sync: every consumer thread of every CTA of the cluster.
// Two alternating mbarriers (count = CS) per CTA; thread r arrives on CTA r's barrier.
__device__ __forceinline__ void cluster_sync_consumers(uint64_t* cbar, uint32_t& nsync,
                                                       uint32_t CS, int tid) {
    fence_cluster();
    named_bar_sync(1, SA_CONSUMER_THREADS);
    uint64_t* bar = cbar + (nsync & 1);
    if (tid < (int)CS) mbar_arrive_remote(bar, (uint32_t)tid);
    mbar_wait_cluster(bar, (nsync >> 1) & 1);
    ++nsync;
}

__device__ __forceinline__ void consumer_bar() { named_bar_sync(1, SA_CONSUMER_THREADS); }

// out(r, k) = <W[row(r)][0:L], X[k][0:L]>  for r < nrows, k < 8.  Eight lanes per weight row.
// X is an [8][L] fp32 matrix in shared memory.  `emit(r, k, v)` is called by lane k of the group.
template <int L, class RowFn, class EmitFn>
__device__ __forceinline__ void rowdots(const float* __restrict__ W, int nrows, const float* X,
                                        int tid, RowFn rowfn, EmitFn emit) {
    const int grp = tid >> 3, gl = tid & 7;
    for (int r = grp; r < nrows; r += SA_CONSUMER_THREADS / 8) {
        const float4* wr = reinterpret_cast<const float4*>(W + (size_t)rowfn(r) * L);
        float4 wv[L / 32];
#pragma unroll
        for (int i = 0; i < L / 32; ++i) wv[i] = __ldg(wr + gl + 8 * i);
        float acc[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            float a = 0.f;
#pragma unroll
            for (int i = 0; i < L / 32; ++i) {
                float4 x = *reinterpret_cast<const float4*>(X + k * L + 4 * gl + 32 * i);
                a = fmaf(wv[i].x, x.x, a); a = fmaf(wv[i].y, x.y, a);
                a = fmaf(wv[i].z, x.z, a); a = fmaf(wv[i].w, x.w, a);
            }
            acc[k] = a;
        }
        float mine = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            float a = acc[k];
            a += __shfl_xor_sync(0xffffffffu, a, 1);
            a += __shfl_xor_sync(0xffffffffu, a, 2);
            a += __shfl_xor_sync(0xffffffffu, a, 4);
            if (gl == k) mine = a;
        }
        emit(r, gl, mine);
    }
}

// LayerNorm of the K rows of an [8][D] smem matrix into another [8][D] smem matrix.
template <int D>
__device__ __forceinline__ void ln_rows(const float* src, float* dst, const float* __restrict__ g,
                                        const float* __restrict__ b, int K, int warp, int lane) {
    if (warp < K) {
        float v[D / 32];
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < D / 32; ++i) { v[i] = src[warp * D + lane + 32 * i]; s += v[i]; }
#pragma unroll
        for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        const float mu = s * (1.f / D);
        float q = 0.f;
#pragma unroll
        for (int i = 0; i < D / 32; ++i) { float t = v[i] - mu; q = fmaf(t, t, q); }
#pragma unroll
        for (int o = 16; o; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
        const float rstd = rsqrtf(q * (1.f / D) + LN_EPS);
#pragma unroll
        for (int i = 0; i < D / 32; ++i) {
            int c = lane + 32 * i;
            dst[warp * D + c] = (v[i] - mu) * rstd * __ldg(g + c) + __ldg(b + c);
        }
    }
}

// ----------------------------------------------------------------------------
// the kernel
// ----------------------------------------------------------------------------
template <int C, int D, int DM>
__global__ void __launch_bounds__(SA_THREADS, 1) sa_forward_kernel(const SAParams p) {
    static_assert(C % 32 == 0 && D % 32 == 0 && DM % 32 == 0, "dims");
    constexpr int KS = C / 16;          // k-steps (logits) == channel blocks (aggregation)
    constexpr int ROWB = C * 2;         // slab row bytes (fp16)
    constexpr int TILE_BYTES = SA_TILE_PX * C * 4;

    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char* slab = smem + p.lay.slab;
    unsigned char* ring = smem + p.lay.ring;
    float* red = reinterpret_cast<float*>(smem + p.lay.red);
    float* uhat = red;                         // [8][C]   (aliases h1)
    float* h1 = red;                           // [8][DM]
    float* sprime = red + 8 * DM;              // [8][D]
    float* lnbuf = red + 8 * DM + 8 * D;       // [8][D]
    float* rs_buf = reinterpret_cast<float*>(smem + p.lay.rs_buf);       // [CS][8][Cc]
    __half* qf_hi = reinterpret_cast<__half*>(smem + p.lay.qfrag);       // [8][C]
    __half* qf_lo = qf_hi + 8 * C;                                       // [8][C]
    float* s_cur = reinterpret_cast<float*>(smem + p.lay.s_cur);         // [8][D]
    float* gi = reinterpret_cast<float*>(smem + p.lay.gates);            // [3*Dc][8]
    float* gh = gi + 3 * 16 * 8;                                         // [3*Dc][8]
    float* colsum_buf = reinterpret_cast<float*>(smem + p.lay.colsum_buf);  // [CS][8]
    float* colsum_w = reinterpret_cast<float*>(smem + p.lay.colsum_w);      // [8 warps][8]
    float* rs_x = reinterpret_cast<float*>(smem + p.lay.rs_x);              // [CS][Cc]
    float* xs_part = reinterpret_cast<float*>(smem + p.lay.xs_part);        // [C]
    float* lng = reinterpret_cast<float*>(smem + p.lay.lnw);                // [C] gamma
    float* lnb = lng + C;                                                   // [C] beta
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + p.lay.bars);
    uint64_t* full = bars;                 // [nstage]
    uint64_t* empty = bars + 8;            // [nstage]
    uint64_t* cbar = bars + 16;            // [2]

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t crank = cluster_ctarank();
    const uint32_t CS = cluster_nctarank();
    const int cid = (int)cluster_id_x(), ncl = (int)cluster_nid_x();
    const int nstage = p.nstage;
    const int rows_cta = p.rows_cta;
    const int ntiles = rows_cta / SA_TILE_PX;
    const int nblk = rows_cta / 128;     // 16-row blocks per warp
    const int px0 = (int)crank * rows_cta;
    const int N = p.N, K = p.K;
    const int Cc = C / (int)CS, Dc = D / (int)CS, Mc = DM / (int)CS;

    // ---------------- one-time setup ----------------
    if (tid == 0) {
        for (int s = 0; s < nstage; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], SA_CONSUMER_WARPS); }
        mbar_init(&cbar[0], CS); mbar_init(&cbar[1], CS);
        fence_mbar_init();
    }
    for (int i = tid; i < C; i += SA_THREADS) { lng[i] = p.ln_in_w[i]; lnb[i] = p.ln_in_b[i]; }
    for (int i = tid; i < 8 * C; i += SA_THREADS) {
        qf_hi[i] = __float2half(0.f); qf_lo[i] = __float2half(0.f);
    }
    for (int i = tid; i < 8 * D; i += SA_THREADS) s_cur[i] = 0.f;
    __syncthreads();
    cluster_barrier_all();   // barriers initialised cluster-wide before any remote arrive

    // ---------------- producer warp: stream feature tiles with TMA bulk copies ----------------
    if (warp == SA_CONSUMER_WARPS) {
        if (lane == 0) {
            const uint64_t pol = l2_policy_evict_first();
            uint32_t n = 0;
            for (int f = cid; f < p.B; f += ncl) {
                const float* src = p.feats + (size_t)f * p.feat_bstride;
                for (int t = 0; t < ntiles; ++t, ++n) {
                    const int st = n % nstage;
                    mbar_wait(&empty[st], ((n / nstage) & 1) ^ 1);
                    const int px = px0 + t * SA_TILE_PX;
                    int nvalid = N - px;
                    nvalid = nvalid < 0 ? 0 : (nvalid > SA_TILE_PX ? SA_TILE_PX : nvalid);
                    if (nvalid > 0) {
                        const uint32_t bytes = (uint32_t)nvalid * C * 4;
                        mbar_arrive_expect_tx(&full[st], bytes);
                        bulk_g2s(ring + (size_t)st * TILE_BYTES, src + (size_t)px * C, bytes, &full[st], pol);
                    } else {
                        mbar_arrive(&full[st]);
                    }
                }
            }
        }
        return;
    }

    // ---------------- consumer warps ----------------
    const int g = lane >> 2, t4 = lane & 3;      // mma fragment coordinates
    const int pxi = lane >> 3, ch8 = lane & 7;   // LayerNorm mapping: 4 pixels x 8 lanes
    const uint32_t slab_u32 = smem_u32(slab);
    uint32_t nsync = 0;
    uint32_t ntile_g = 0;
    int pidx = 0;
    const bool do_prof = (p.prof != nullptr) && cid == 0 && crank == 0 && tid == 0;
#define SA_STAMP() do { if (do_prof && pidx < p.prof_cap) p.prof[pidx++] = globaltimer_ns(); } while (0)

    for (int f = cid; f < p.B; f += ncl) {
        // ---- initial slots -> s_cur (every CTA keeps the full K x D state) ----
        for (int i = tid; i < K * D; i += SA_CONSUMER_THREADS)
            s_cur[i] = __ldg(p.slots_in + (size_t)f * K * D + i);
        consumer_bar();

        SA_STAMP();   // 0: frame start (slots loaded)
        float xs_own = 0.f;   // reduced sum_n x^[n][c] for the channel this thread owns (tid < 8*Cc)

        for (int it = 0; it < p.n_iter; ++it) {
            // ================= step D: q~ slice = LNq(S) W_qk, all-gather as fp16 hi/lo ========
            {
                ln_rows<D>(s_cur, lnbuf, p.ln_q_w, p.ln_q_b, K, warp, lane);
                consumer_bar();
                rowdots<D>(p.w_qk, Cc, lnbuf, tid,
                           [&](int r) { return (int)crank * Cc + r; },
                           [&](int r, int k, float v) {
                               if (k < K) {
                                   const int c = (int)crank * Cc + r;
                                   const __half hi = __float2half_rn(v);
                                   const __half lo = __float2half_rn(v - __half2float(hi));
                                   const uint32_t a_hi = smem_u32(qf_hi + k * C + c);
                                   const uint32_t a_lo = smem_u32(qf_lo + k * C + c);
                                   for (uint32_t rk = 0; rk < CS; ++rk) {
                                       st_cluster_u16(mapa(a_hi, rk), __half_as_ushort(hi));
                                       st_cluster_u16(mapa(a_lo, rk), __half_as_ushort(lo));
                                   }
                               }
                           });
                cluster_sync_consumers(cbar, nsync, CS, tid);
            }
            SA_STAMP();   // 1: q~ all-gathered

            // ================= attention pass over this CTA's pixels =================
            uint32_t bq_hi[KS][2], bq_lo[KS][2];
#pragma unroll
            for (int ks = 0; ks < KS; ++ks) {
                const int c0 = 16 * ks + 2 * t4;
                bq_hi[ks][0] = *reinterpret_cast<const uint32_t*>(qf_hi + g * C + c0);
                bq_hi[ks][1] = *reinterpret_cast<const uint32_t*>(qf_hi + g * C + c0 + 8);
                bq_lo[ks][0] = *reinterpret_cast<const uint32_t*>(qf_lo + g * C + c0);
                bq_lo[ks][1] = *reinterpret_cast<const uint32_t*>(qf_lo + g * C + c0 + 8);
            }
            float uacc[KS][4];
#pragma unroll
            for (int cb = 0; cb < KS; ++cb) { uacc[cb][0] = uacc[cb][1] = uacc[cb][2] = uacc[cb][3] = 0.f; }
            float cs0 = 0.f, cs1 = 0.f;            // column sums for slots 2*t4, 2*t4+1
            float xs[C / 32][4];                   // sum_n x^ partials (first pass only)
#pragma unroll
            for (int i = 0; i < C / 32; ++i) { xs[i][0] = xs[i][1] = xs[i][2] = xs[i][3] = 0.f; }
            const bool want_mask = (p.seg_mask != nullptr) && (it == p.n_iter - 1);

            for (int j = 0; j < nblk; ++j) {
                const int rb = (j * SA_CONSUMER_WARPS + warp) * 16;   // slab row base of this block
                if (it == 0) {
                    // ---- stream 4 tiles: LayerNorm 4 pixels per warp per tile into the slab ----
                    for (int q = 0; q < 4; ++q, ++ntile_g) {
                        const int st = ntile_g % nstage;
                        mbar_wait(&full[st], (ntile_g / nstage) & 1);
                        const int tile = 4 * j + q;
                        const int px = px0 + tile * SA_TILE_PX + warp * 4 + pxi;
                        const bool valid = px < N;
                        const float* tp = reinterpret_cast<const float*>(ring + (size_t)st * TILE_BYTES) +
                                          (warp * 4 + pxi) * C + 4 * ch8;
                        float4 v[C / 32];
#pragma unroll
                        for (int i = 0; i < C / 32; ++i)
                            v[i] = valid ? *reinterpret_cast<const float4*>(tp + 32 * i)
                                         : make_float4(0.f, 0.f, 0.f, 0.f);
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&empty[st]);
                        float s = 0.f;
#pragma unroll
                        for (int i = 0; i < C / 32; ++i) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
                        s += __shfl_xor_sync(0xffffffffu, s, 1);
                        s += __shfl_xor_sync(0xffffffffu, s, 2);
                        s += __shfl_xor_sync(0xffffffffu, s, 4);
                        const float mu = s * (1.f / C);
                        float qv = 0.f;
#pragma unroll
                        for (int i = 0; i < C / 32; ++i) {
                            v[i].x -= mu; v[i].y -= mu; v[i].z -= mu; v[i].w -= mu;
                            qv = fmaf(v[i].x, v[i].x, qv); qv = fmaf(v[i].y, v[i].y, qv);
                            qv = fmaf(v[i].z, v[i].z, qv); qv = fmaf(v[i].w, v[i].w, qv);
                        }
                        qv += __shfl_xor_sync(0xffffffffu, qv, 1);
                        qv += __shfl_xor_sync(0xffffffffu, qv, 2);
                        qv += __shfl_xor_sync(0xffffffffu, qv, 4);
                        const float rstd = valid ? rsqrtf(qv * (1.f / C) + LN_EPS) : 0.f;
                        const int row = rb + q * 4 + pxi;
                        unsigned char* rowp = slab + (size_t)row * ROWB + (ch8 & 1) * 8;
#pragma unroll
                        for (int i = 0; i < C / 32; ++i) {
                            const float4 gm = *reinterpret_cast<const float4*>(lng + 4 * ch8 + 32 * i);
                            const float4 bt = *reinterpret_cast<const float4*>(lnb + 4 * ch8 + 32 * i);
                            float y0 = valid ? fmaf(v[i].x * rstd, gm.x, bt.x) : 0.f;
                            float y1 = valid ? fmaf(v[i].y * rstd, gm.y, bt.y) : 0.f;
                            float y2 = valid ? fmaf(v[i].z * rstd, gm.z, bt.z) : 0.f;
                            float y3 = valid ? fmaf(v[i].w * rstd, gm.w, bt.w) : 0.f;
                            xs[i][0] += y0; xs[i][1] += y1; xs[i][2] += y2; xs[i][3] += y3;
                            const int chunk = ((ch8 >> 1) + 4 * i) ^ (row & 7);
                            uint2 pk; pk.x = pack_h2(y0, y1); pk.y = pack_h2(y2, y3);
                            *reinterpret_cast<uint2*>(rowp + chunk * 16) = pk;
                        }
                    }
                    __syncwarp();
                }

                // ---- logits for 16 pixels x 8 slots (log2 domain; scale folded into q~) ----
                float lg[4] = {0.f, 0.f, 0.f, 0.f};
                {
                    const int row = rb + (lane & 7) + ((lane >> 3) & 1) * 8;
                    const uint32_t rowa = slab_u32 + row * ROWB;
#pragma unroll
                    for (int ks = 0; ks < KS; ++ks) {
                        uint32_t a[4];
                        ldsm_x4(a, rowa + (((2 * ks + (lane >> 4)) ^ (row & 7)) << 4));
                        mma_f16(lg, a, bq_hi[ks][0], bq_hi[ks][1]);
                        mma_f16(lg, a, bq_lo[ks][0], bq_lo[ks][1]);
                    }
                }
                // pixel ids of fragment rows g and g+8
                const int pxa = px0 + (4 * j + (g >> 2)) * SA_TILE_PX + warp * 4 + (g & 3);
                const int pxb = pxa + 2 * SA_TILE_PX;
                const bool s0ok = (2 * t4) < K, s1ok = (2 * t4 + 1) < K;
                float pa0, pa1, pb0, pb1;
                {
                    float m = fmaxf(s0ok ? lg[0] : -INFINITY, s1ok ? lg[1] : -INFINITY);
                    m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 1));
                    m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 2));
                    float e0 = s0ok ? exp2f(lg[0] - m) : 0.f, e1 = s1ok ? exp2f(lg[1] - m) : 0.f;
                    float su = e0 + e1;
                    su += __shfl_xor_sync(0xffffffffu, su, 1);
                    su += __shfl_xor_sync(0xffffffffu, su, 2);
                    const float inv = (pxa < N) ? __fdividef(1.f, su) : 0.f;
                    pa0 = e0 * inv; pa1 = e1 * inv;
                    m = fmaxf(s0ok ? lg[2] : -INFINITY, s1ok ? lg[3] : -INFINITY);
                    m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 1));
                    m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 2));
                    e0 = s0ok ? exp2f(lg[2] - m) : 0.f; e1 = s1ok ? exp2f(lg[3] - m) : 0.f;
                    su = e0 + e1;
                    su += __shfl_xor_sync(0xffffffffu, su, 1);
                    su += __shfl_xor_sync(0xffffffffu, su, 2);
                    const float invb = (pxb < N) ? __fdividef(1.f, su) : 0.f;
                    pb0 = e0 * invb; pb1 = e1 * invb;
                }
                if (want_mask) {
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
                // probabilities as fp16(1024*a); column sums from the ROUNDED values
                const __half2 ha = __floats2half2_rn(pa0 * SA_PSCALE, pa1 * SA_PSCALE);
                const __half2 hb = __floats2half2_rn(pb0 * SA_PSCALE, pb1 * SA_PSCALE);
                {
                    const float2 fa = __half22float2(ha), fb = __half22float2(hb);
                    cs0 += fa.x + fb.x; cs1 += fa.y + fb.y;
                }
                const uint32_t b0 = movmatrix_t(*reinterpret_cast<const uint32_t*>(&ha));
                const uint32_t b1 = movmatrix_t(*reinterpret_cast<const uint32_t*>(&hb));
                // ---- aggregation U^T[16 ch x 8 slots] += x^T[16 ch x 16 px] * P[16 px x 8 slots] ----
                {
                    const int row = rb + (lane & 7) + (lane >> 4) * 8;
                    const uint32_t rowa = slab_u32 + row * ROWB;
#pragma unroll
                    for (int cb = 0; cb < KS; ++cb) {
                        uint32_t a[4];
                        ldsm_x4_t(a, rowa + (((2 * cb + ((lane >> 3) & 1)) ^ (row & 7)) << 4));
                        mma_f16(uacc[cb], a, b0, b1);
                    }
                }
            }

            SA_STAMP();   // 2: pass done
            // ================= CTA-level reduction of the pass partials =================
            // column sums: reduce over g (lanes sharing t4), then one row per warp
            cs0 += __shfl_xor_sync(0xffffffffu, cs0, 4);  cs1 += __shfl_xor_sync(0xffffffffu, cs1, 4);
            cs0 += __shfl_xor_sync(0xffffffffu, cs0, 8);  cs1 += __shfl_xor_sync(0xffffffffu, cs1, 8);
            cs0 += __shfl_xor_sync(0xffffffffu, cs0, 16); cs1 += __shfl_xor_sync(0xffffffffu, cs1, 16);
            if (lane < 4) { colsum_w[warp * 8 + 2 * lane] = cs0; colsum_w[warp * 8 + 2 * lane + 1] = cs1; }
            if (it == 0) {
                // sum_n x^: reduce over the 4 pixel groups of the warp, then over warps via `red`
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
                consumer_bar();
                if (tid < C) {
                    float a = 0.f;
#pragma unroll
                    for (int w8 = 0; w8 < SA_CONSUMER_WARPS; ++w8) a += red[w8 * C + tid];
                    xs_part[tid] = a;
                }
                consumer_bar();
            }
            // tree reduction of uacc over the 8 warps: slots of `red` are [4][KS*4][32] floats
            {
                constexpr int NREG = KS * 4;
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
                if (warp >= 4) put(warp - 4);
                consumer_bar();
                if (warp < 4) add(warp);
                if (warp == 2 || warp == 3) put(warp);
                consumer_bar();
                if (warp < 2) add(2 + warp);
                if (warp == 1) put(1);
                consumer_bar();
                if (warp == 0) add(1);
            }
            SA_STAMP();   // 3: CTA tree reduction done
            // ================= E1: reduce-scatter U^T (by channel owner) + colsum / xsum =========
            if (warp == 0) {
                const uint32_t rs_u32 = smem_u32(rs_buf);
#pragma unroll
                for (int cb = 0; cb < KS; ++cb)
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const int c = 16 * cb + g + 8 * (e >> 1);
                        const int slot = 2 * t4 + (e & 1);
                        const int owner = c / Cc, cl = c % Cc;
                        st_cluster_f32(mapa(rs_u32 + (((int)crank * 8 + slot) * Cc + cl) * 4, owner), uacc[cb][e]);
                    }
            } else if (warp == 1) {
                // colsum of this CTA -> every CTA's colsum_buf[crank][slot]
                if (lane < 8) {
                    float a = 0.f;
#pragma unroll
                    for (int w8 = 0; w8 < SA_CONSUMER_WARPS; ++w8) a += colsum_w[w8 * 8 + lane];
                    const uint32_t a_cs = smem_u32(colsum_buf + crank * 8 + lane);
                    for (uint32_t rk = 0; rk < CS; ++rk) st_cluster_f32(mapa(a_cs, rk), a);
                }
            } else if (warp == 2 && it == 0) {
                const uint32_t rx_u32 = smem_u32(rs_x);
                for (int c = lane; c < C; c += 32) {
                    const int owner = c / Cc, cl = c % Cc;
                    st_cluster_f32(mapa(rx_u32 + ((int)crank * Cc + cl) * 4, owner), xs_part[c]);
                }
            }
            cluster_sync_consumers(cbar, nsync, CS, tid);

            SA_STAMP();   // 4: E1 synced
            // ================= owner reduce -> u^ slice, E2: all-gather u^ =================
            if (tid < 8 * Cc) {
                const int slot = tid / Cc, cl = tid % Cc;
                if (it == 0) {
                    float a = 0.f;
                    for (uint32_t rk = 0; rk < CS; ++rk) a += rs_x[rk * Cc + cl];
                    xs_own = a;
                }
                if (slot < K) {
                    float us = 0.f, cs = 0.f;
                    for (uint32_t rk = 0; rk < CS; ++rk) {
                        us += rs_buf[(rk * 8 + slot) * Cc + cl];
                        cs += colsum_buf[rk * 8 + slot];
                    }
                    const float num = us * (1.f / SA_PSCALE) + p.eps * xs_own;
                    const float den = cs * (1.f / SA_PSCALE) + (float)N * p.eps;
                    const float u = num / den;
                    const uint32_t a_u = smem_u32(uhat + slot * C + (int)crank * Cc + cl);
                    for (uint32_t rk = 0; rk < CS; ++rk) st_cluster_f32(mapa(a_u, rk), u);
                }
            }
            cluster_sync_consumers(cbar, nsync, CS, tid);

            SA_STAMP();   // 5: E2 synced (u^ everywhere)
            // ================= step A: GRU gates for this CTA's Dc columns =================
            rowdots<C>(p.w_iv, 3 * Dc, uhat, tid,
                       [&](int r) { return (r / Dc) * D + (int)crank * Dc + (r % Dc); },
                       [&](int r, int k, float v) { gi[r * 8 + k] = v; });
            rowdots<D>(p.w_hh, 3 * Dc, s_cur, tid,
                       [&](int r) { return (r / Dc) * D + (int)crank * Dc + (r % Dc); },
                       [&](int r, int k, float v) { gh[r * 8 + k] = v; });
            consumer_bar();
            if (tid < K * Dc) {
                const int k = tid / Dc, jj = tid % Dc, jc = (int)crank * Dc + jj;
                const float r_ = sigmoidf_(gi[jj * 8 + k] + __ldg(p.b_ih + jc) + gh[jj * 8 + k] + __ldg(p.b_hh + jc));
                const float z_ = sigmoidf_(gi[(Dc + jj) * 8 + k] + __ldg(p.b_ih + D + jc) +
                                           gh[(Dc + jj) * 8 + k] + __ldg(p.b_hh + D + jc));
                const float n_ = tanhf(gi[(2 * Dc + jj) * 8 + k] + __ldg(p.b_ih + 2 * D + jc) +
                                       r_ * (gh[(2 * Dc + jj) * 8 + k] + __ldg(p.b_hh + 2 * D + jc)));
                const float sp = (1.f - z_) * n_ + z_ * s_cur[k * D + jc];
                const uint32_t a_s = smem_u32(sprime + k * D + jc);
                for (uint32_t rk = 0; rk < CS; ++rk) st_cluster_f32(mapa(a_s, rk), sp);
            }
            cluster_sync_consumers(cbar, nsync, CS, tid);

            SA_STAMP();   // 6: step A + E3
            // ================= step B: hidden slice of the residual MLP =================
            ln_rows<D>(sprime, lnbuf, p.ln_m_w, p.ln_m_b, K, warp, lane);
            consumer_bar();
            rowdots<D>(p.w1, Mc, lnbuf, tid,
                       [&](int r) { return (int)crank * Mc + r; },
                       [&](int r, int k, float v) {
                           if (k < K) {
                               const int col = (int)crank * Mc + r;
                               const float h = fmaxf(v + __ldg(p.b1 + col), 0.f);
                               const uint32_t a_h = smem_u32(h1 + k * DM + col);
                               for (uint32_t rk = 0; rk < CS; ++rk) st_cluster_f32(mapa(a_h, rk), h);
                           }
                       });
            cluster_sync_consumers(cbar, nsync, CS, tid);

            SA_STAMP();   // 7: step B + E4
            // ================= step C: new slots slice =================
            const bool last = (it == p.n_iter - 1);
            rowdots<DM>(p.w2, Dc, h1, tid,
                        [&](int r) { return (int)crank * Dc + r; },
                        [&](int r, int k, float v) {
                            if (k < K) {
                                const int col = (int)crank * Dc + r;
                                const float sn = sprime[k * D + col] + v + __ldg(p.b2 + col);
                                if (last) {
                                    p.slots_out[((size_t)f * K + k) * D + col] = sn;
                                } else {
                                    const uint32_t a_s = smem_u32(s_cur + k * D + col);
                                    for (uint32_t rk = 0; rk < CS; ++rk) st_cluster_f32(mapa(a_s, rk), sn);
                                }
                            }
                        });
            SA_STAMP();   // 8: step C done (before E5 sync)
            if (last) {
                // peers may still read sprime / h1 of this iteration; the next frame's first
                // remote writes (q~) go to qfrag only, and its first cluster sync orders the rest.
                consumer_bar();
                break;
            }
            cluster_sync_consumers(cbar, nsync, CS, tid);
        }
    }
    cluster_sync_consumers(cbar, nsync, CS, tid);   // nobody leaves while peers may still signal it
}

// ----------------------------------------------------------------------------
// host side
// ----------------------------------------------------------------------------
static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

int sa_plan(int N, int C, int D, int DM, int cluster_size, int smem_limit, SAPlan* plan) {
    if (!((C == 128 && D == 128 && DM == 256) || (C == 192 && D == 192 && DM == 384))) return -1;
    int cs = cluster_size;
    if (cs == 0) cs = (C == 128) ? 8 : 16;
    if (cs != 8 && cs != 16) return -1;
    if (C % cs || D % cs || DM % cs) return -1;
    int rows = (N + cs - 1) / cs;
    rows = (rows + 127) / 128 * 128;
    SALayout L;
    size_t off = 0;
    L.slab = (uint32_t)off; off += (size_t)rows * C * 2; off = align_up(off, 128);
    const size_t red_bytes = (size_t)C * 128;             // 4 slots x (C/4 regs) x 32 lanes x 4 B
    if (red_bytes < (size_t)(8 * DM + 16 * D) * 4) return -1;
    L.red = (uint32_t)off; off += red_bytes;
    L.rs_buf = (uint32_t)off; off += (size_t)8 * C * 4;
    L.qfrag = (uint32_t)off; off += (size_t)8 * C * 2 * 2;
    L.s_cur = (uint32_t)off; off += (size_t)8 * D * 4;
    L.gates = (uint32_t)off; off += (size_t)2 * 3 * 16 * 8 * 4;
    L.colsum_buf = (uint32_t)off; off += (size_t)cs * 8 * 4;
    L.colsum_w = (uint32_t)off; off += (size_t)8 * 8 * 4;
    L.rs_x = (uint32_t)off; off += (size_t)C * 4;
    L.xs_part = (uint32_t)off; off += (size_t)C * 4;
    L.lnw = (uint32_t)off; off += (size_t)2 * C * 4;
    L.bars = (uint32_t)off; off += 18 * 8;
    off = align_up(off, 128);
    L.ring = (uint32_t)off;
    const size_t tile_bytes = (size_t)SA_TILE_PX * C * 4;
    if (off + 2 * tile_bytes > (size_t)smem_limit) return -1;
    int nstage = (int)(((size_t)smem_limit - off) / tile_bytes);
    if (nstage > 8) nstage = 8;
    off += (size_t)nstage * tile_bytes;
    plan->lay = L;
    plan->cluster_size = cs;
    plan->rows_cta = rows;
    plan->nstage = nstage;
    plan->smem_bytes = off;
    return 0;
}

template <int C, int D, int DM>
static cudaError_t config_t(const SAPlan& plan, cudaLaunchConfig_t* cfg, cudaLaunchAttribute* attr, int* max_clusters) {
    auto kern = sa_forward_kernel<C, D, DM>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan.smem_bytes);
    if (e != cudaSuccess) return e;
    if (plan.cluster_size > 8) {
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
        if (e != cudaSuccess) return e;
    }
    *cfg = cudaLaunchConfig_t{};
    cfg->blockDim = dim3(SA_THREADS);
    cfg->dynamicSmemBytes = plan.smem_bytes;
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = plan.cluster_size;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg->attrs = attr;
    cfg->numAttrs = 1;
    // persistent: as many clusters as can be co-resident (cached per configuration)
    static int cached_clusters[2][2] = {{0, 0}, {0, 0}};
    int& nc = cached_clusters[C == 128 ? 0 : 1][plan.cluster_size == 8 ? 0 : 1];
    if (nc == 0) {
        cfg->gridDim = dim3(plan.cluster_size * 8);
        int n = 0;
        e = cudaOccupancyMaxActiveClusters(&n, kern, cfg);
        if (e != cudaSuccess) return e;
        if (n < 1) return cudaErrorLaunchOutOfResources;
        nc = n;
    }
    *max_clusters = nc;
    return cudaSuccess;
}

template <int C, int D, int DM>
static cudaError_t launch_t(const SAParams& p, const SAPlan& plan, int max_clusters_hint, cudaStream_t st) {
    cudaLaunchConfig_t cfg;
    cudaLaunchAttribute attr[1];
    int ncl = 0;
    cudaError_t e = config_t<C, D, DM>(plan, &cfg, attr, &ncl);
    if (e != cudaSuccess) return e;
    cfg.stream = st;
    if (max_clusters_hint > 0 && ncl > max_clusters_hint) ncl = max_clusters_hint;
    if (ncl > p.B) ncl = p.B;
    cfg.gridDim = dim3(ncl * plan.cluster_size);
    return cudaLaunchKernelEx(&cfg, sa_forward_kernel<C, D, DM>, p);
}

cudaError_t sa_launch(const SAParams& p, const SAPlan& plan, int C, int max_clusters_hint, cudaStream_t st) {
    if (C == 128) return launch_t<128, 128, 256>(p, plan, max_clusters_hint, st);
    return launch_t<192, 192, 384>(p, plan, max_clusters_hint, st);
}

int sa_max_clusters(int C, int cluster_size) {
    int dev = 0, smem = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return -1;
    if (cudaDeviceGetAttribute(&smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev) != cudaSuccess) return -1;
    SAPlan plan;
    const int D = C, DM = 2 * C;
    if (sa_plan(4096, C, D, DM, cluster_size, smem, &plan)) return -1;
    cudaLaunchConfig_t cfg;
    cudaLaunchAttribute attr[1];
    int ncl = 0;
    cudaError_t e = (C == 128) ? config_t<128, 128, 256>(plan, &cfg, attr, &ncl)
                               : config_t<192, 192, 384>(plan, &cfg, attr, &ncl);
    return e == cudaSuccess ? ncl : -1;
}

cudaError_t sa_fold_launch(const float* wq, const float* wk, const float* wv, const float* w_ih,
                           float* w_qk, float* w_iv, int C, int D, cudaStream_t st) {
    const int total = C * D + 3 * D * C;
    const float qscale = (1.0f / sqrtf((float)D)) * 1.4426950408889634f;
    sa_fold_kernel<<<(total + 255) / 256, 256, 0, st>>>(wq, wk, wv, w_ih, w_qk, w_iv, C, D, qscale);
    return cudaGetLastError();
}

}  // namespace sfb
