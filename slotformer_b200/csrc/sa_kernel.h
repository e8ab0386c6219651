// Internal interface between the C ABI (capi.cu) and the Slot Attention kernels.
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stddef.h>

namespace sfb {

// slot-update weights as packed fp16 hi/lo panel pairs (built once per call by sa_prep):
// pair (nb, kb) of a [N][Kd] matrix = 64x64 hi panel then 64x64 lo panel, 128B-swizzled rows
struct SAWeightsDev {
    const __half* w_qk;   // [C][D]   scale*log2e * (Wq^T Wk)^T : q~ = LNq(S) W_qk^T
    const __half* w_iv;   // [3D][C]  W_ih Wv
    const __half* w_hh;   // [3D][D]
    const __half* w1;     // [Dm][D]
    const __half* w2;     // [D][Dm]
    const float *b_ih, *b_hh, *b1, *b2;
    const float *ln_q_w, *ln_q_b, *ln_m_w, *ln_m_b;
    const float *ln_in_w, *ln_in_b;   // norm_inputs affine (applied here, not in the pass)
    const float* wbeta;               // [D]  sum_c beta_c W_qk[c][:]  -> per-slot logit bias
};

// workspace carve-up (byte offsets), computed by sa_workspace_layout
struct SAWorkspace {
    size_t w_qk, w_iv, w_hh, w1, w2;   // packed hi/lo panel pairs
    size_t wbeta;      // [D] fp32
    size_t qt;         // [B] x { hi [8][C] fp16, lo [8][C] fp16, logit bias [8] fp32 }
    int qt_stride;     // halves per frame
    size_t partials;   // [B][nchunk][pstride] fp32
    size_t xsum;       // [B][C] fp32
    size_t xhat;       // [xhat_frames][N16/16][16*C] fp16 (swizzled 16-pixel tiles)
    size_t total;
    int nchunk, chunk_px, pstride, n16, xhat_frames;
};

struct SAPassParams {
    const void* feats;         // fp32 or bf16 (feat_esize bytes per element)
    int feat_esize;
    long long feat_bstride;    // elements between frames
    __half* xhat;              // nullable when there is no later pass
    long long xhat_fstride;    // bytes between frames of the x^ ring / tile buffer (tcgen05 passes)
    int write_xsum;            // tcgen05 ring pass: also write sum_n t[n] (the first iteration runs on given tiles)
    const __half* qt;          // per frame: hi [8][C], lo [8][C], 8 fp32 logit biases
    float* partials;
    float* seg_mask;           // nullable; written only by the last pass
    const float *ln_w, *ln_b;  // norm_inputs
    int B, N, K, nchunk, chunk_px, pstride, n16, xhat_frames;
    int frame0;                // first frame of this launch (chunked scheduling)
    int nframes;               // frames in this launch
    unsigned long long* prof;
    int prof_cap;
    int cta_limited;           // the grid is capped below one CTA per SM (batch pipeline): the passes are SM-bound
    int split;                 // mma.sync first pass on warp pairs: 1 / 0 = forced on / off, -1 = when the grid is capped
    int dbg;                   // debug switches (debug build, SFB_DBG env): 1 = no proxy fence, 2 = no x^ store, 4 = skip LN+MMA work
    int xhat_keep;             // SFB_SA_XHAT_KEEP: x^ stores stay evict_last even under a capped grid
    int reverse;               // tcgen05 passes: walk the items last-to-first (odd iterations: the pass starts on the x^
                               // tiles the previous pass touched last, which are still in L2)
};

struct SAUpdateParams {
    SAWeightsDev w;
    const float* partials;
    float* xsum;               // [B][C]; written when first != 0
    const float* slots_prev;   // [B][K][D] state before this update (slots_in or slots_out)
    float* slots_out;          // [B][K][D]
    __half* qt;                // see SAWorkspace::qt; written when do_q != 0
    int qt_stride;
    int B, N, K, nchunk, pstride;
    int frame0, nframes;
    int do_update, do_q, first;
    int qt_swz;                // q~ written as the swizzled tcgen05 operand image (sa_pass_tc.cu) instead of [hi | lo] rows
    float eps;
};

void sa_workspace_layout(int B, int chunk_frames, int N, int C, int D, int DM, int n_iter, SAWorkspace* ws);
cudaError_t sa_prep_launch(const float* wq, const float* wk, const float* wv, const float* w_ih,
                           const float* w_hh, const float* w1, const float* w2, const float* ln_in_w,
                           const float* ln_in_b, char* ws_base,
                           const SAWorkspace& ws, int C, int D, int DM, cudaStream_t st);
cudaError_t sa_pass_launch(const SAPassParams& p, int C, bool first, int sms, int smem_limit, cudaStream_t st);
// first pass with LayerNorm and tensor-core halves of a tile on different warps (sa_pass_split.cu)
bool sa_pass_split_supported(const SAPassParams& p, int C);
cudaError_t sa_pass_split_launch(const SAPassParams& p, int sms, cudaStream_t st);
// tcgen05 passes (sa_pass_tc.cu): C = 128; q~ must be in the swizzled operand layout (SAUpdateParams::qt_swz)
bool sa_pass_tc_supported(const SAPassParams& p, int C);
cudaError_t sa_pass_tc_launch(const SAPassParams& p, bool first, int sms, cudaStream_t st);
cudaError_t sa_update_launch(const SAUpdateParams& p, int C, int sms, cudaStream_t st, bool capped = false);
// encoder tail (enc_tail.cu)
size_t enc_tail_workspace_bytes();
cudaError_t enc_tail_prep_launch(const float* pos_w, const float* pos_b, const float* ln_w, const float* ln_b,
                                 const float* w1, const float* b1, const float* w2, const float* b2, char* ws,
                                 cudaStream_t st);
cudaError_t enc_tail_launch(const float* cnn, long long frame_stride, int frames, int H, int W, void* tiles,
                            const char* ws, int sms, int tiles_frame, bool nhwc, cudaStream_t st);
bool sa_shape_supported(int C, int D, int DM);

}  // namespace sfb
