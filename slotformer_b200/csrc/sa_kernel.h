// Internal interface between the C ABI (capi.cu) and the Slot Attention kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>

namespace sfb {

// byte offsets of the shared-memory regions of sa_forward_kernel (computed on the host)
struct SALayout {
    uint32_t slab, ring, red, rs_buf, qfrag, s_cur, gates, colsum_buf, colsum_w, rs_x, xs_part,
        lnw, bars;
};

struct SAPlan {
    SALayout lay;
    int cluster_size;
    int rows_cta;   // pixels owned by one CTA (multiple of 128)
    int nstage;     // TMA ring depth
    size_t smem_bytes;
};

struct SAParams {
    const float* feats;
    long long feat_bstride;   // elements between consecutive frames
    const float* slots_in;
    float* slots_out;
    float* seg_mask;          // nullable
    const float *ln_in_w, *ln_in_b, *ln_q_w, *ln_q_b;
    const float* w_qk;        // [C][D]  folded, includes scale*log2(e)
    const float* w_iv;        // [3D][C] folded W_ih Wv
    const float *w_hh, *b_ih, *b_hh;
    const float *ln_m_w, *ln_m_b, *w1, *b1, *w2, *b2;
    int B, N, K, n_iter;
    float eps;
    int rows_cta, nstage;
    SALayout lay;
    unsigned long long* prof;   // optional timeline buffer (debug), cluster 0 / CTA 0 / thread 0
    int prof_cap;
};

int sa_plan(int N, int C, int D, int DM, int cluster_size, int smem_limit, SAPlan* plan);
cudaError_t sa_launch(const SAParams& p, const SAPlan& plan, int C, int max_clusters_hint, cudaStream_t st);
int sa_max_clusters(int C, int cluster_size);
cudaError_t sa_fold_launch(const float* wq, const float* wk, const float* wv, const float* w_ih,
                           float* w_qk, float* w_iv, int C, int D, cudaStream_t st);

}  // namespace sfb
