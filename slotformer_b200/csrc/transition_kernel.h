// Internal interface between the C ABI (capi.cu) and the slot-transition kernel (transition.cu).
#pragma once
#include <cuda_runtime.h>

namespace sfb {

static constexpr int TR_THREADS = 512;
static constexpr int TR_MAX_WIDTH = 1024;     // widest activation (3D, ffn, 4H, 2D, D + H)
static constexpr int TR_NBUF = 4;             // activation buffers [width][8 rows] fp32 in shared memory
static constexpr int TR_MAX_OPS = 64;
static constexpr int TR_MAX_VEC = 6144;       // biases + LayerNorm affine of the whole program, staged in shared memory

enum TrCode { TR_END = 0, TR_LOAD, TR_STORE, TR_LN, TR_LINEAR, TR_ATTN, TR_LSTM, TR_SAMPLE };
static constexpr int TR_F_RELU = 1;
static constexpr int TR_F_ADD = 2;            // LINEAR: dst += result (residual)
static constexpr int TR_F_NOSYNC = 4;         // LOAD / STORE: the next step does not depend on this one (no barrier)

struct TrOp { int code; int a[7]; };

struct TrParams {
    const float* blob;          // packed weights (sfb_transition_prepare)
    const float* in[4];         // 0 previous slots / initial latents, 1 h_in, 2 c_in, 3 noise   (nullptr = zeros / absent)
    float* out[4];              // 0 dist [B,K,2D], 1 initial slots [B,K,D], 2 h_out, 3 c_out    (nullptr = not wanted)
    int B, K, nops;
    int vec_floats;             // the blob starts with this many floats of bias / LayerNorm vectors
    unsigned long long* prof;   // debug build: per-step globaltimer stamps of CTA 0 (nullptr = off)
    int prof_cap;
    TrOp ops[TR_MAX_OPS];
};

cudaError_t transition_launch(const TrParams& p, int cluster_size, cudaStream_t st);
int transition_max_clusters(int K, int cluster_size);
cudaError_t tr_pack_weight(const float* W, float* dst, int N, int Kd, int krow0, cudaStream_t st);
cudaError_t tr_pack_vec(const float* a, const float* b, float* dst, int n, cudaStream_t st);

}  // namespace sfb
