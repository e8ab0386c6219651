// SAVi slot transition t -> t+1 as ONE kernel.
//
// Reference: the per-frame glue between two Slot Attention calls in StoSAVi.encode
// (base_slots/models/savi.py:393-410):
//     latents = predictor(prev_slots)          predictor.py:20-44 (Transformer over the K slots of a clip) or
//                                              predictor.py:47-74 (LayerNorm + residual MLP), optionally wrapped
//                                              in predictor.py:76-113 (LSTM cell over time + out_projector)
//     dist    = kernel_dist_layer(latents)     savi.py:200-212 (Linear [-> LayerNorm -> ReLU -> Linear])
//     slots0  = mu (+ noise * exp(logvar / 2)) savi.py:355-363
// In PyTorch that is ~40 launches per frame on [B*K, 128] activations: pure launch latency.  Here one thread-block
// CLUSTER owns one clip: the K <= 8 slot rows live in shared memory (feature-major, [feature][8 rows] fp32), the kernel
// interprets a short program of LOAD / LN / LINEAR / ATTN / LSTM / SAMPLE / STORE steps, and every LINEAR is split over the
// CTAs of the cluster by output feature: each CTA streams its slice of the (pre-transposed, k-major) fp32 weight matrix
// from L2 with 128-bit loads, keeps 8 rows x 4 features of fp32 accumulators per thread, and writes its slice of the
// result into the activation buffer of EVERY CTA of the cluster through distributed shared memory; one cluster barrier
// per LINEAR is the only inter-CTA synchronisation.  The cheap steps (LayerNorm, the K x K attention, the LSTM gates)
// are computed redundantly by every CTA.  All arithmetic is fp32 (FFMA): the reference path is fp32 and the step is
// latency-bound, not FLOP-bound (0.9 M weights x K rows per clip).
#include "common.cuh"
#include "transition_kernel.h"

namespace sfb {

static constexpr int TR_BUF = TR_MAX_WIDTH * 8;            // floats per activation buffer
static constexpr int TR_RED_FLOATS = (TR_THREADS + 16) * 32;   // k-slice partial sums (float4 pitch NG + 1)
static constexpr size_t TR_SMEM_BYTES = (size_t)(TR_NBUF * TR_BUF + TR_RED_FLOATS + 1024 + TR_MAX_VEC) * sizeof(float);

__device__ __forceinline__ void st_cluster_f4(uint32_t local_addr, uint32_t rank, float4 v) {
    const uint32_t ra = mapa(local_addr, rank);
    asm volatile("st.shared::cluster.v4.f32 [%0], {%1, %2, %3, %4};" :: "r"(ra), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// dst[n][r] (+)= act(sum_k src[k][r] * Wt[k][n] + bias[n]) for this CTA's slice of the N output features
template <int ROWS>
__device__ __forceinline__ void tr_linear(const TrOp& op, float* bufs, float* red, const float* __restrict__ blob,
                                          const float* vec, int tid, int crank, int csize) {
    const float* src = bufs + op.a[0] * TR_BUF;
    float* dst = bufs + op.a[1] * TR_BUF;
    // a: src, dst, Kd | ksplit << 16, N, weight offset, bias offset, flags | src2 << 8 | src2 feature offset << 12
    // (rows k >= ksplit of the operand come from buffer src2: the LSTM gates read [x ; h] from two buffers)
    const int Kd = op.a[2] & 0xffff, N = op.a[3], flags = op.a[6];
    const int ks = (op.a[2] >> 16) ? (op.a[2] >> 16) : Kd;
    const float* src2 = bufs + ((flags >> 8) & 3) * TR_BUF + ((flags >> 12) & 0xfff) * 8 - ks * 8;   // indexed by k
    const float* __restrict__ Wt = blob + op.a[4];
    const int Nc = N / csize, n0 = crank * Nc, NG = Nc >> 2;        // float4 feature groups of this CTA
    // k slices: every thread busy when the slice of features is narrow, at least 8 k per slice, and the partial
    // sums S x 8 x (NG + 1) float4 within the scratch (TR_RED_FLOATS)
    int S = TR_THREADS / NG;
    if (S > 64) S = 64;
    if (S > (Kd >> 3)) S = Kd >> 3;
    if (S > 528 / (NG + 1)) S = 528 / (NG + 1);
    if (S < 1) S = 1;
    const int Kc = (Kd + S - 1) / S;
    const int g = tid % NG, s = tid / NG;
    float4* red4 = reinterpret_cast<float4*>(red);
    const int P = NG + 1;                                               // float4 pitch of a (slice, half, j) row
    if (s < S) {
        float acc[ROWS][4];
#pragma unroll
        for (int r = 0; r < ROWS; ++r) { acc[r][0] = acc[r][1] = acc[r][2] = acc[r][3] = 0.f; }
        const int k0 = s * Kc, k1 = (k0 + Kc) < Kd ? (k0 + Kc) : Kd;
        const float4* wp = reinterpret_cast<const float4*>(Wt + (size_t)k0 * N + n0 + 4 * g);
        const int wstride = N >> 2;
        int k = k0;
        for (; k + 8 <= k1; k += 8) {           // 8 independent 128-bit loads in flight per thread
            float4 w[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) w[u] = __ldg(wp + (size_t)u * wstride);
            wp += 8 * (size_t)wstride;
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const float* xp = ((k + u) < ks ? src : src2) + (k + u) * 8;
                const float4 xa = *reinterpret_cast<const float4*>(xp);
                float x[8] = {xa.x, xa.y, xa.z, xa.w, 0.f, 0.f, 0.f, 0.f};
                if (ROWS > 4) {
                    const float4 xb = *reinterpret_cast<const float4*>(xp + 4);
                    x[4] = xb.x; x[5] = xb.y; x[6] = xb.z; x[7] = xb.w;
                }
#pragma unroll
                for (int r = 0; r < ROWS; ++r) {
                    acc[r][0] = fmaf(x[r], w[u].x, acc[r][0]);
                    acc[r][1] = fmaf(x[r], w[u].y, acc[r][1]);
                    acc[r][2] = fmaf(x[r], w[u].z, acc[r][2]);
                    acc[r][3] = fmaf(x[r], w[u].w, acc[r][3]);
                }
            }
        }
        for (; k < k1; ++k) {
            const float4 w = __ldg(wp);
            wp += wstride;
            const float* xp = (k < ks ? src : src2) + k * 8;
            const float4 xa = *reinterpret_cast<const float4*>(xp);
            float x[8] = {xa.x, xa.y, xa.z, xa.w, 0.f, 0.f, 0.f, 0.f};
            if (ROWS > 4) {
                const float4 xb = *reinterpret_cast<const float4*>(xp + 4);
                x[4] = xb.x; x[5] = xb.y; x[6] = xb.z; x[7] = xb.w;
            }
#pragma unroll
            for (int r = 0; r < ROWS; ++r) {
                acc[r][0] = fmaf(x[r], w.x, acc[r][0]);
                acc[r][1] = fmaf(x[r], w.y, acc[r][1]);
                acc[r][2] = fmaf(x[r], w.z, acc[r][2]);
                acc[r][3] = fmaf(x[r], w.w, acc[r][3]);
            }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float v[8];
#pragma unroll
            for (int r = 0; r < 8; ++r) v[r] = r < ROWS ? acc[r < ROWS ? r : 0][j] : 0.f;
            red4[((s * 2 + 0) * 4 + j) * P + g] = make_float4(v[0], v[1], v[2], v[3]);
            red4[((s * 2 + 1) * 4 + j) * P + g] = make_float4(v[4], v[5], v[6], v[7]);
        }
    }
    __syncthreads();
    // slice sums + bias (+ ReLU) (+ residual) -> this feature slice in the buffer of every CTA of the cluster
    for (int idx = tid; idx < Nc * 2; idx += TR_THREADS) {
        const int n = idx >> 1, half = idx & 1, gg = n >> 2, j = n & 3;
        float4 a = red4[((0 * 2 + half) * 4 + j) * P + gg];
        for (int ss = 1; ss < S; ++ss) {
            const float4 b = red4[((ss * 2 + half) * 4 + j) * P + gg];
            a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
        }
        if (op.a[5] >= 0) {
            const float bi = vec[op.a[5] + n0 + n];
            a.x += bi; a.y += bi; a.z += bi; a.w += bi;
        }
        if (flags & TR_F_RELU) { a.x = fmaxf(a.x, 0.f); a.y = fmaxf(a.y, 0.f); a.z = fmaxf(a.z, 0.f); a.w = fmaxf(a.w, 0.f); }
        float* dp = dst + (n0 + n) * 8 + half * 4;
        if (flags & TR_F_ADD) {
            const float4 o = *reinterpret_cast<const float4*>(dp);
            a.x += o.x; a.y += o.y; a.z += o.z; a.w += o.w;
        }
        if (csize == 1) {
            *reinterpret_cast<float4*>(dp) = a;
        } else {
            const uint32_t la = smem_u32(dp);
            for (int q = 0; q < csize; ++q) st_cluster_f4(la, (uint32_t)q, a);
        }
    }
}

// LayerNorm over `width` features of every row (biased variance, eps 1e-5, affine; optional ReLU); in place allowed.
// Two warps per row (features split in halves, partial sums exchanged through `scratch`); three sweeps over shared memory
// instead of a register-resident row keep the code small: every step of this kernel runs only a few times per launch.
__device__ __forceinline__ void tr_layernorm(const TrOp& op, float* bufs, const float* vec, float* scratch, int K, int warp, int lane) {
    const float* src = bufs + op.a[0] * TR_BUF;
    float* dst = bufs + op.a[1] * TR_BUF;
    const int W = op.a[2];
    const float* gam = vec + op.a[3];
    const float* bet = vec + op.a[4];
    const bool relu = (op.a[5] & TR_F_RELU) != 0;
    const int r = warp & 7, hf = warp >> 3;                 // 16 warps: row, half of the features
    const int f0 = hf * (W >> 1), f1 = f0 + (W >> 1);
    float sum = 0.f;
    for (int f = f0 + lane; f < f1; f += 32) sum += src[f * 8 + r];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    if (lane == 0) scratch[warp] = sum;
    __syncthreads();
    const float mean = (scratch[r] + scratch[r + 8]) / (float)W;
    float sq = 0.f;
    for (int f = f0 + lane; f < f1; f += 32) { const float d = src[f * 8 + r] - mean; sq += d * d; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
    if (lane == 0) scratch[16 + warp] = sq;
    __syncthreads();
    const float rstd = 1.f / sqrtf((scratch[16 + r] + scratch[24 + r]) / (float)W + 1e-5f);
    for (int f = f0 + lane; f < f1; f += 32) {
        float y = (src[f * 8 + r] - mean) * rstd * gam[f] + bet[f];
        if (relu) y = fmaxf(y, 0.f);
        dst[f * 8 + r] = r < K ? y : 0.f;
    }
}

// multi-head self-attention over the K rows of the clip: src = [q | k | v] (3D features), dst = heads concatenated
__device__ __forceinline__ void tr_attention(const TrOp& op, float* bufs, float* scratch, int K, int tid) {
    const float* src = bufs + op.a[0] * TR_BUF;
    float* dst = bufs + op.a[1] * TR_BUF;
    const int D = op.a[2], H = op.a[3], dh = D / H;
    const float scale = 1.f / sqrtf((float)dh);
    for (int t = tid; t < H * 64; t += TR_THREADS) {
        const int h = t >> 6, i = (t >> 3) & 7, j = t & 7;
        float s = 0.f;
        const float* q = src + (h * dh) * 8 + i;
        const float* k = src + (D + h * dh) * 8 + j;
        for (int c = 0; c < dh; ++c) s = fmaf(q[c * 8] * scale, k[c * 8], s);
        scratch[t] = (i < K && j < K) ? s : -INFINITY;
    }
    __syncthreads();
    for (int t = tid; t < H * 8; t += TR_THREADS) {
        float* row = scratch + t * 8;
        const int i = t & 7;
        if (i < K) {
            float m = row[0];
            for (int j = 1; j < K; ++j) m = fmaxf(m, row[j]);
            float e[8], sum = 0.f;
            for (int j = 0; j < 8; ++j) { e[j] = j < K ? expf(row[j] - m) : 0.f; sum += e[j]; }
            const float inv = 1.f / sum;
            for (int j = 0; j < 8; ++j) row[j] = e[j] * inv;
        } else {
            for (int j = 0; j < 8; ++j) row[j] = 0.f;
        }
    }
    __syncthreads();
    for (int t = tid; t < D * 8; t += TR_THREADS) {
        const int f = t >> 3, i = t & 7, h = f / dh;
        const float* p = scratch + (h * 8 + i) * 8;
        const float* v = src + (2 * D + f) * 8;
        float o = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) o = fmaf(p[j], v[j], o);
        dst[t] = o;
    }
}

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

template <int ROWS>
__global__ void __launch_bounds__(TR_THREADS, 1) transition_kernel(const TrParams p) {
    extern __shared__ __align__(16) float tr_smem[];
    float* bufs = tr_smem;
    float* red = bufs + TR_NBUF * TR_BUF;
    float* scratch = red + TR_RED_FLOATS;
    float* vec = scratch + 1024;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int csize = (int)cluster_nctarank(), crank = (int)cluster_ctarank();
    const int b = blockIdx.x / csize;
    const int K = p.K;

    // the program itself: one sweep into shared memory, so that the step records are not fetched from the kernel-parameter
    // (constant) bank with a run-time index step by step (measured: no difference in the step times either way)
    __shared__ TrOp s_ops[TR_MAX_OPS];
    for (int i = tid; i < p.nops * 8; i += TR_THREADS)
        reinterpret_cast<int*>(s_ops)[i] = reinterpret_cast<const int*>(p.ops)[i];
    // every bias / LayerNorm vector of the program: one coalesced sweep into shared memory (no global-load latency
    // inside the steps)
    for (int i = tid * 4; i < p.vec_floats; i += TR_THREADS * 4)
        *reinterpret_cast<float4*>(vec + i) = __ldg(reinterpret_cast<const float4*>(p.blob + i));
    // every CTA of the cluster has started before any peer writes into its shared memory (racecheck: "block that might
    // not have entered yet" on the first LINEAR without this)
    if (csize > 1) cluster_barrier_all(); else __syncthreads();
    const bool do_prof = p.prof != nullptr && blockIdx.x == 0 && tid == 0;
    for (int ip = 0; ip < p.nops; ++ip) {
        const TrOp op = s_ops[ip];
        bool cluster_wide = false, nosync = false;
        if (do_prof && ip < p.prof_cap) p.prof[ip] = (unsigned long long)clock64();
        switch (op.code) {
        case TR_LOAD: {         // a: buf, input index, width, row stride, feature offset in buf, clip stride (-1: K * row stride)
            float* dst = bufs + op.a[0] * TR_BUF + op.a[4] * 8;
            const float* g = p.in[op.a[1]];
            const long long bs = op.a[5] >= 0 ? (long long)op.a[5] : (long long)K * op.a[3];
            const int W = op.a[2];
            for (int t = tid; t < W * 8; t += TR_THREADS) {
                const int r = t / W, f = t - r * W;         // coalesced global reads, strided smem writes
                dst[f * 8 + r] = (g != nullptr && r < K) ? __ldg(g + (size_t)b * bs + (size_t)r * op.a[3] + f) : 0.f;
            }
            nosync = (op.a[6] & TR_F_NOSYNC) != 0;
            break;
        }
        case TR_STORE: {        // a: buf, output index, width, row stride, feature offset in buf
            const float* src = bufs + op.a[0] * TR_BUF + op.a[4] * 8;
            float* g = p.out[op.a[1]];
            const int W = op.a[2];
            if (g != nullptr && crank == 0)
                for (int t = tid; t < W * K; t += TR_THREADS) {
                    const int r = t / W, f = t - r * W;
                    g[((size_t)b * K + r) * op.a[3] + f] = src[f * 8 + r];
                }
            nosync = (op.a[6] & TR_F_NOSYNC) != 0;
            break;
        }
        case TR_LN:
            tr_layernorm(op, bufs, vec, scratch, K, warp, lane);
            break;
        case TR_LINEAR:
            tr_linear<ROWS>(op, bufs, red, p.blob, vec, tid, crank, csize);
            cluster_wide = csize > 1;
            break;
        case TR_ATTN:
            tr_attention(op, bufs, scratch, K, tid);
            break;
        case TR_LSTM: {         // a: gates buf [i | f | g | o], c buf, h' dst buf, hidden size  (torch gate order)
            const float* G = bufs + op.a[0] * TR_BUF;
            const float* c = bufs + op.a[1] * TR_BUF;
            float* hd = bufs + op.a[2] * TR_BUF;
            const int H = op.a[3];
            for (int t = tid; t < H * 8; t += TR_THREADS) {
                const int f = t >> 3, r = t & 7;
                const float ig = sigmoidf_(G[t]), fg = sigmoidf_(G[H * 8 + t]);
                const float gg = tanhf(G[2 * H * 8 + t]), og = sigmoidf_(G[3 * H * 8 + t]);
                const float cn = fg * c[t] + ig * gg;
                const float hn = og * tanhf(cn);
                hd[t] = r < K ? hn : 0.f;
                if (crank == 0 && r < K) {
                    if (p.out[2]) p.out[2][((size_t)b * K + r) * H + f] = hn;
                    if (p.out[3]) p.out[3][((size_t)b * K + r) * H + f] = cn;
                }
            }
            break;
        }
        case TR_SAMPLE: {       // a: dist buf [mu | log var], D:  slots = mu (+ noise * exp(log var / 2))
            const float* d = bufs + op.a[0] * TR_BUF;
            const int D = op.a[1];
            const float* nz = p.in[3];
            if (p.out[1] != nullptr && crank == 0)
                for (int t = tid; t < D * K; t += TR_THREADS) {
                    const int r = t / D, f = t - r * D;
                    float v = d[f * 8 + r];
                    const size_t gi = ((size_t)b * K + r) * D + f;
                    if (nz != nullptr) v += __ldg(nz + gi) * expf(0.5f * d[(D + f) * 8 + r]);
                    p.out[1][gi] = v;
                }
            break;
        }
        default: break;
        }
        if (cluster_wide) cluster_barrier_all(); else if (!nosync) __syncthreads();
    }
    if (do_prof && p.nops < p.prof_cap) p.prof[p.nops] = (unsigned long long)clock64();
    if (csize > 1) cluster_barrier_all();      // no CTA exits while a peer may still write into its shared memory
}

// W [N][Kd] (PyTorch Linear layout) -> blob rows [krow0 + k][n] of a k-major matrix with N columns
__global__ void tr_pack_weight_kernel(const float* __restrict__ W, float* __restrict__ dst, int N, int Kd, int krow0) {
    __shared__ float tile[32][33];
    const int n0 = blockIdx.x * 32, k0 = blockIdx.y * 32;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int n = n0 + i, k = k0 + threadIdx.x;
        tile[i][threadIdx.x] = (n < N && k < Kd) ? W[(size_t)n * Kd + k] : 0.f;
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int k = k0 + i, n = n0 + threadIdx.x;
        if (n < N && k < Kd) dst[(size_t)(krow0 + k) * N + n] = tile[threadIdx.x][i];
    }
}
__global__ void tr_pack_vec_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ dst, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = a[i] + (b ? b[i] : 0.f);
}

cudaError_t tr_pack_weight(const float* W, float* dst, int N, int Kd, int krow0, cudaStream_t st) {
    dim3 grid((N + 31) / 32, (Kd + 31) / 32), block(32, 8);
    tr_pack_weight_kernel<<<grid, block, 0, st>>>(W, dst, N, Kd, krow0);
    return cudaGetLastError();
}
cudaError_t tr_pack_vec(const float* a, const float* b, float* dst, int n, cudaStream_t st) {
    tr_pack_vec_kernel<<<(n + 255) / 256, 256, 0, st>>>(a, b, dst, n);
    return cudaGetLastError();
}

template <int ROWS>
static cudaError_t tr_launch_t(const TrParams& p, int csize, cudaStream_t st) {
    auto kern = transition_kernel<ROWS>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TR_SMEM_BYTES);
    if (e != cudaSuccess) return e;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(p.B * csize), 1, 1);
    cfg.blockDim = dim3(TR_THREADS, 1, 1);
    cfg.dynamicSmemBytes = TR_SMEM_BYTES;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)csize;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, p);
}

template <int ROWS>
static int tr_max_clusters_t(int csize) {
    auto kern = transition_kernel<ROWS>;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TR_SMEM_BYTES) != cudaSuccess) return 0;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)csize, 1, 1);
    cfg.blockDim = dim3(TR_THREADS, 1, 1);
    cfg.dynamicSmemBytes = TR_SMEM_BYTES;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)csize;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

// clusters of `csize` CTAs that can be resident at once (a GPC holds only whole clusters: 8-CTA clusters fill 16 of its
// 18-20 SMs); depends on the device only
int transition_max_clusters(int K, int csize) {
    if (K <= 4) return tr_max_clusters_t<4>(csize);
    if (K <= 6) return tr_max_clusters_t<6>(csize);
    return tr_max_clusters_t<8>(csize);
}

cudaError_t transition_launch(const TrParams& p, int csize, cudaStream_t st) {
    if (p.K <= 4) return tr_launch_t<4>(p, csize, st);
    if (p.K <= 6) return tr_launch_t<6>(p, csize, st);
    return tr_launch_t<8>(p, csize, st);
}

}  // namespace sfb
