// Internal interface between the C ABI (capi.cu) and the rollout kernels.
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stddef.h>

namespace sfb {

static constexpr int RO_MAX_LAYERS = 16;

struct ROLayer {
    const __half *wqkv, *wo, *w1, *w2;                    // packed fp16 panels (workspace)
    const float *bqkv, *bo, *b1, *b2, *ln1w, *ln1b, *ln2w, *ln2b;   // fp32 originals
};

struct ROParams {
    const float* hist;   // [B][T_h*K][Ds]
    float* pred;         // [B][pred_len*K][Ds]
    const __half* w_in;  // packed [d][Ds]
    const float* b_in;
    const __half* w_out; // packed [Ds][d]
    const __half *w_in_lo, *w_out_lo;   // fp16(w - fp16(w)) of the two, same packing (engine B: three-term products)
    const float* b_out;
    const float* pe;     // [pe_tokens][d]
    const float* par_g;  // [layers][par_floats]: bqkv | bo | b1 | b2 | ln1w | ln1b | ln2w | ln2b (workspace, fp32)
    int B, hist_tokens, K, Ds, d, F, heads, layers, pred_len, mode, cond_tokens, pe_tokens;
    int hg;              // heads per attention group
    int fc;              // FFN hidden chunk (columns)
    int lmax;            // max window tokens over the rollout
    int lda, ldb;        // row strides (halves) of the fp16 activation buffers
    int par_floats;      // per-layer parameter block (biases + LayerNorm) staged in smem
    int par_double;      // 1: double-buffered (next layer prefetched), 0: single buffer (big shapes)
    int nstage;          // weight-panel ring depth
    int stage_tiles;     // engine B: 16 KB weight tiles per ring stage (1 or 2)
    uint32_t off_h, off_a, off_b, off_bars, off_ring, off_par;   // shared-memory byte offsets
    unsigned long long* prof;       // optional timeline buffer (debug)
    int prof_cap;
    int dbg;             // SFB_DBG switches (engine B): 1 = no weight copies, 2 = no MMAs, 4 = no epilogue math
    ROLayer layer[RO_MAX_LAYERS];
};

// fp32 [N][Kd] -> fp16 128x64 weight tiles (pairs of swizzled 64x64 panels), rows padded to 128
cudaError_t ro_pack2_launch(const float* src, __half* dst, int N, int Kd, cudaStream_t st, bool lo = false);
cudaError_t umma_test_launch(const __half* Wp, const float* X, float* out, int M, int N, int K, cudaStream_t st);
// chooses hg / fc / buffer layout; returns 0 or -1 if the shape cannot be kept on chip
// engine A (mma.sync from a TMA-fed panel ring; any supported shape)
int ro_mma_plan(ROParams* p, int smem_limit, size_t* smem_bytes);
cudaError_t ro_mma_launch(const ROParams& p, size_t smem_bytes, cudaStream_t st);
// engine B (tcgen05 + TMEM, swap-AB; windows up to 64 tokens that fit on chip)
int ro_umma_plan(ROParams* p, int smem_limit, size_t* smem_bytes);
cudaError_t ro_umma_launch(const ROParams& p, size_t smem_bytes, cudaStream_t st);

}  // namespace sfb
