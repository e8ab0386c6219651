// C ABI of libsfb200.so (declared in include/sfb200.h).  Plain pointers and sizes only;
// validation + launch; no host synchronisation, no persistent allocations.
#include "../../include/sfb200.h"
#include "sa_kernel.h"
#include "ro_kernel.h"
#include "decode_kernel.h"
#include "transition_kernel.h"

#include <atomic>
#include <cstring>
#include <cstdlib>

namespace {

std::atomic<long long> g_launches{0};    // statistics only (sfb_launch_count); no call reads it
#ifdef SFB_DEBUG
// Debug build only (python -m slotformer_b200.build --debug -> libsfb200_debug.so): timeline buffer and the
// SFB_DBG kernel switches.  The product library has neither: its calls depend on their arguments alone.
unsigned long long* g_prof = nullptr;   // device buffer, see sfb_debug_set_profile
int g_prof_cap = 0;
int debug_switches() { const char* dv = getenv("SFB_DBG"); return dv ? atoi(dv) : 0; }
#else
constexpr unsigned long long* g_prof = nullptr;
constexpr int g_prof_cap = 0;
constexpr int debug_switches() { return 0; }
#endif

inline int cuda_err(cudaError_t e) { return e == cudaSuccess ? SFB_OK : (SFB_E_CUDA_BASE - (int)e); }
inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

struct DevInfo { int ok; int smem_optin; int sms; int cc; };

// per-device properties, looked up once per device
int device_info(DevInfo* out) {
    static DevInfo cache[64];
    static std::atomic<int> ready[64];
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return cuda_err(e);
    if (dev < 0 || dev >= 64) return SFB_E_UNSUPPORTED_ARCH;
    if (!ready[dev].load(std::memory_order_acquire)) {
        DevInfo d{};
        int major = 0, minor = 0;
        if ((e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev)) != cudaSuccess) return cuda_err(e);
        if ((e = cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev)) != cudaSuccess) return cuda_err(e);
        if ((e = cudaDeviceGetAttribute(&d.smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev)) != cudaSuccess) return cuda_err(e);
        if ((e = cudaDeviceGetAttribute(&d.sms, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return cuda_err(e);
        d.cc = major * 10 + minor;
        d.ok = 1;
        cache[dev] = d;
        ready[dev].store(1, std::memory_order_release);
    }
    *out = cache[dev];
    return SFB_OK;
}


}  // namespace

extern "C" {

int sfb_version(void) { return SFB_VERSION; }

#ifdef SFB_DEBUG
void sfb_debug_set_profile(void* device_buf, int capacity) {
    g_prof = reinterpret_cast<unsigned long long*>(device_buf);
    g_prof_cap = device_buf ? capacity : 0;
}
#endif

long long sfb_launch_count(void) { return g_launches.load(); }

const char* sfb_strerror(int code) {
    switch (code) {
        case SFB_OK: return "ok";
        case SFB_E_BAD_SHAPE: return "unsupported or inconsistent shape";
        case SFB_E_BAD_ALIGN: return "pointer or stride is not 16-byte aligned";
        case SFB_E_UNSUPPORTED_ARCH: return "device is not sm_100 (B200)";
        case SFB_E_WORKSPACE: return "workspace too small";
        case SFB_E_NULL: return "required pointer is NULL";
        default: break;
    }
    if (code <= SFB_E_CUDA_BASE) return cudaGetErrorString((cudaError_t)(SFB_E_CUDA_BASE - code));
    return "unknown error";
}

#ifdef SFB_DEBUG
int sfb_debug_umma_gemm(const float* W, const float* X, float* out, int M, int N, int K, void* workspace,
                        size_t workspace_bytes, void* stream) {
    if (!W || !X || !out || !workspace) return SFB_E_NULL;
    if (M < 128 || M % 128 || N < 16 || N > 128 || N % 16 || K < 64 || K % 64) return SFB_E_BAD_SHAPE;
    if (workspace_bytes < (size_t)M * K * 2) return SFB_E_WORKSPACE;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    cudaError_t e = sfb::ro_pack2_launch(W, reinterpret_cast<__half*>(workspace), M, K, st);
    if (e != cudaSuccess) return cuda_err(e);
    e = sfb::umma_test_launch(reinterpret_cast<const __half*>(workspace), X, out, M, N, K, st);
    return cuda_err(e);
}
#endif

// ------------------------------------------------------------------------------------------
// Slot Attention
// ------------------------------------------------------------------------------------------
static int sa_pick_chunk(int B, int N, int C, int n_iter, int chunk_frames) {
    (void)N; (void)C; (void)n_iter;
    // Default: the whole batch is one scheduling chunk (passes and updates are full-GPU launches).
    // Smaller chunks keep the fp16 x^ ring L2 resident but serialise the update kernels; they pay
    // off only once chunks are overlapped on several streams ("next").
    if (chunk_frames > 0) return chunk_frames < B ? chunk_frames : B;
    return B;
}

size_t sfb_sa_workspace_bytes(int B, int N, int C, int D, int Dm, int n_iter, int chunk_frames) {
    if (B <= 0 || N <= 0 || C <= 0 || D <= 0 || Dm <= 0 || n_iter <= 0) return 0;
    sfb::SAWorkspace ws;
    const int chunk = sa_pick_chunk(B, N, C, n_iter, chunk_frames);
    sfb::sa_workspace_layout(B, chunk, N, C, D, Dm, n_iter, &ws);
    return ws.total;
}

int sfb_sa_prepare(const sfb_sa_weights* w, int C, int D, int Dm, void* workspace, size_t workspace_bytes,
                   void* stream) {
    if (!w || !workspace) return SFB_E_NULL;
    const float* const* wp = reinterpret_cast<const float* const*>(w);
    for (size_t i = 0; i < sizeof(sfb_sa_weights) / sizeof(const float*); ++i)
        if (!wp[i]) return SFB_E_NULL;
    if (!sfb::sa_shape_supported(C, D, Dm)) return SFB_E_BAD_SHAPE;
    if (!aligned16(workspace)) return SFB_E_BAD_ALIGN;
    sfb::SAWorkspace ws;
    sfb::sa_workspace_layout(1, 1, 16, C, D, Dm, 1, &ws);      // the weight regions come first
    if (workspace_bytes < ws.qt) return SFB_E_WORKSPACE;
    cudaError_t e = sfb::sa_prep_launch(w->project_q_1_weight, w->project_k_weight, w->project_v_weight,
                                        w->gru_weight_ih, w->gru_weight_hh, w->mlp_1_weight, w->mlp_3_weight,
                                        w->norm_inputs_weight, w->norm_inputs_bias,
                                        reinterpret_cast<char*>(workspace), ws, C, D, Dm,
                                        reinterpret_cast<cudaStream_t>(stream));
    if (e != cudaSuccess) return cuda_err(e);
    g_launches.fetch_add(2);
    return SFB_OK;
}

int sfb_sa_forward(const void* feats, int feat_dtype, int64_t feat_batch_stride,
                   const float* slots_in, float* slots_out, float* seg_mask,
                   const sfb_sa_weights* w, int B, int N, int C, int D, int Dm, int K,
                   int n_iter, float eps, int chunk_frames, int max_ctas, unsigned int flags,
                   void* workspace, size_t workspace_bytes, void* stream) {
    if (B == 0) return SFB_OK;
    if (!feats || !slots_in || !slots_out || !w || !workspace) return SFB_E_NULL;
    const float* const* wp = reinterpret_cast<const float* const*>(w);
    for (size_t i = 0; i < sizeof(sfb_sa_weights) / sizeof(const float*); ++i)
        if (!wp[i]) return SFB_E_NULL;
    if (B < 0 || N < 1 || K < 1 || K > 8 || n_iter < 1 || max_ctas < 0) return SFB_E_BAD_SHAPE;
    if (feat_dtype != SFB_DTYPE_F32 && feat_dtype != SFB_DTYPE_BF16 && feat_dtype != SFB_DTYPE_TILES16) return SFB_E_BAD_SHAPE;
    if (!sfb::sa_shape_supported(C, D, Dm)) return SFB_E_BAD_SHAPE;
    const bool tiles_in = feat_dtype == SFB_DTYPE_TILES16;
    if (tiles_in && (C != 128 || (flags & SFB_SA_NO_TCGEN05))) return SFB_E_BAD_SHAPE;   // tiles feed the tcgen05 passes only
    if (tiles_in ? (feat_batch_stride < (int64_t)sfb_enc_tail_tiles_bytes(1, N, C) / 2)
                 : (feat_batch_stride < (int64_t)N * C)) return SFB_E_BAD_SHAPE;
    if (feat_batch_stride & 7) return SFB_E_BAD_ALIGN;
    if (!aligned16(feats) || !aligned16(slots_in) || !aligned16(slots_out) || !aligned16(workspace))
        return SFB_E_BAD_ALIGN;
    // (given tiles are read in place: no x^ ring of its own, as with a single iteration)
    if (workspace_bytes < sfb_sa_workspace_bytes(B, N, C, D, Dm, tiles_in ? 1 : n_iter, chunk_frames)) return SFB_E_WORKSPACE;
    DevInfo di;
    int rc = device_info(&di);
    if (rc) return rc;
    if (di.cc / 10 != 10) return SFB_E_UNSUPPORTED_ARCH;

    // persistent streaming passes: one CTA per SM, or fewer when the caller shares the GPU with the rollout
    bool cta_limited = false;
    if (max_ctas > 0 && max_ctas < di.sms) { di.sms = max_ctas; cta_limited = true; }
    const int chunk = sa_pick_chunk(B, N, C, n_iter, chunk_frames);
    sfb::SAWorkspace ws;
    sfb::sa_workspace_layout(B, chunk, N, C, D, Dm, tiles_in ? 1 : n_iter, &ws);
    char* base = reinterpret_cast<char*>(workspace);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    cudaError_t e;

    auto H = [&](size_t off) { return reinterpret_cast<const __half*>(base + off); };
    sfb::SAUpdateParams up{};
    up.w.w_qk = H(ws.w_qk); up.w.w_iv = H(ws.w_iv); up.w.w_hh = H(ws.w_hh); up.w.w1 = H(ws.w1); up.w.w2 = H(ws.w2);
    up.w.b_ih = w->gru_bias_ih; up.w.b_hh = w->gru_bias_hh; up.w.b1 = w->mlp_1_bias; up.w.b2 = w->mlp_3_bias;
    up.w.ln_q_w = w->project_q_0_weight; up.w.ln_q_b = w->project_q_0_bias;
    up.w.ln_m_w = w->mlp_0_weight; up.w.ln_m_b = w->mlp_0_bias;
    up.w.ln_in_w = w->norm_inputs_weight; up.w.ln_in_b = w->norm_inputs_bias;
    up.w.wbeta = reinterpret_cast<const float*>(base + ws.wbeta);
    up.qt_stride = ws.qt_stride;
    up.partials = reinterpret_cast<const float*>(base + ws.partials);
    up.xsum = reinterpret_cast<float*>(base + ws.xsum);
    up.slots_out = slots_out;
    up.qt = reinterpret_cast<__half*>(base + ws.qt);
    up.B = B; up.N = N; up.K = K; up.nchunk = ws.nchunk; up.pstride = ws.pstride; up.eps = eps;

    sfb::SAPassParams pp{};
    pp.feats = feats;
    pp.feat_esize = (feat_dtype == SFB_DTYPE_BF16) ? 2 : 4;
    pp.feat_bstride = feat_batch_stride;
    pp.qt = up.qt;
    pp.partials = reinterpret_cast<float*>(base + ws.partials);
    pp.ln_w = w->norm_inputs_weight; pp.ln_b = w->norm_inputs_bias;
    pp.B = B; pp.N = N; pp.K = K; pp.nchunk = ws.nchunk; pp.chunk_px = ws.chunk_px;
    pp.pstride = ws.pstride; pp.n16 = ws.n16; pp.xhat_frames = ws.xhat_frames > 0 ? ws.xhat_frames : 1;
    pp.xhat_fstride = (long long)ws.n16 * C * 2;
    if (tiles_in) { pp.xhat_frames = B; pp.xhat_fstride = (long long)feat_batch_stride * 2; }
    pp.prof = g_prof; pp.prof_cap = g_prof_cap;
    pp.cta_limited = cta_limited ? 1 : 0;
    pp.split = (flags & SFB_SA_SPLIT_ON) ? 1 : ((flags & SFB_SA_SPLIT_OFF) ? 0 : -1);
    pp.dbg = debug_switches();
    pp.xhat_keep = (flags & SFB_SA_XHAT_KEEP) ? 1 : 0;
    // tcgen05 passes (C = 128) unless the caller asks for the mma.sync ones; q~ then lives in the operand layout
    const bool use_tc = !(flags & SFB_SA_NO_TCGEN05) && sfb::sa_pass_tc_supported(pp, C);
    up.qt_swz = use_tc ? 1 : 0;
    if (tiles_in && !use_tc) return SFB_E_BAD_SHAPE;

    for (int f0 = 0; f0 < B; f0 += chunk) {
        const int nf = (B - f0) < chunk ? (B - f0) : chunk;
        up.frame0 = f0; up.nframes = nf;
        pp.frame0 = f0; pp.nframes = nf;
        // q~ of the initial slots
        up.do_update = 0; up.do_q = 1; up.first = 0; up.slots_prev = slots_in;
        if ((e = sfb::sa_update_launch(up, C, di.sms, st, cta_limited)) != cudaSuccess) return cuda_err(e);
        g_launches.fetch_add(1);
        for (int it = 0; it < n_iter; ++it) {
            const bool last = (it == n_iter - 1);
            pp.xhat = tiles_in ? reinterpret_cast<__half*>(const_cast<void*>(feats))
                               : ((n_iter > 1) ? reinterpret_cast<__half*>(base + ws.xhat) : nullptr);
            pp.seg_mask = last ? seg_mask : nullptr;
            pp.write_xsum = (tiles_in && it == 0) ? 1 : 0;
            pp.reverse = it & 1;          // alternate the walk: a pass starts where the previous one ended (L2 reuse)
            e = use_tc ? sfb::sa_pass_tc_launch(pp, it == 0 && !tiles_in, di.sms, st)
                       : sfb::sa_pass_launch(pp, C, it == 0, di.sms, di.smem_optin, st);
            if (e != cudaSuccess) return cuda_err(e);
            g_launches.fetch_add(1);
            up.do_update = 1; up.do_q = last ? 0 : 1; up.first = (it == 0);
            up.slots_prev = (it == 0) ? slots_in : slots_out;
            if ((e = sfb::sa_update_launch(up, C, di.sms, st, cta_limited)) != cudaSuccess) return cuda_err(e);
            g_launches.fetch_add(1);
        }
    }
    return SFB_OK;
}

// ------------------------------------------------------------------------------------------
// Encoder tail (section 8 f1)
// ------------------------------------------------------------------------------------------
size_t sfb_enc_tail_workspace_bytes(int C) { return C == 128 ? sfb::enc_tail_workspace_bytes() : 0; }

size_t sfb_enc_tail_tiles_bytes(int frames, int N, int C) {
    if (frames < 0 || N < 1 || C != 128) return 0;
    // the tile count of a frame is the one sfb_sa_forward's items cover (sa_workspace_layout: n16 / 128)
    sfb::SAWorkspace ws;
    sfb::sa_workspace_layout(1, 1, N, C, C, 2 * C, 1, &ws);
    return (size_t)frames * ws.n16 * C * 2;
}

int sfb_enc_tail_prepare(const sfb_enc_tail_weights* w, int C, void* workspace, size_t workspace_bytes, void* stream) {
    if (!w || !workspace) return SFB_E_NULL;
    const float* const* wp = reinterpret_cast<const float* const*>(w);
    for (size_t i = 0; i < sizeof(sfb_enc_tail_weights) / sizeof(const float*); ++i)
        if (!wp[i]) return SFB_E_NULL;
    if (C != 128) return SFB_E_BAD_SHAPE;
    if (!aligned16(workspace)) return SFB_E_BAD_ALIGN;
    if (workspace_bytes < sfb::enc_tail_workspace_bytes()) return SFB_E_WORKSPACE;
    cudaError_t e = sfb::enc_tail_prep_launch(
        w->encoder_pos_embedding_dense_weight, w->encoder_pos_embedding_dense_bias, w->encoder_out_layer_0_weight,
        w->encoder_out_layer_0_bias, w->encoder_out_layer_1_weight, w->encoder_out_layer_1_bias,
        w->encoder_out_layer_3_weight, w->encoder_out_layer_3_bias, reinterpret_cast<char*>(workspace),
        reinterpret_cast<cudaStream_t>(stream));
    if (e != cudaSuccess) return cuda_err(e);
    g_launches.fetch_add(1);
    return SFB_OK;
}

int sfb_enc_tail_forward(const float* cnn_out, int64_t frame_stride, int frames, int H, int W, int C, void* tiles,
                         size_t tiles_bytes, const void* workspace, size_t workspace_bytes, int max_ctas, unsigned int flags,
                         void* stream) {
    if (frames == 0) return SFB_OK;
    if (!cnn_out || !tiles || !workspace) return SFB_E_NULL;
    if (frames < 0 || H < 1 || W < 1 || C != 128 || max_ctas < 0) return SFB_E_BAD_SHAPE;
    const long long N = (long long)H * W;
    if (N > (1 << 24) || (N & 3)) return SFB_E_BAD_SHAPE;                  // TMA row pitch: multiples of 16 bytes
    if (frame_stride < 64 * N || (frame_stride & 3)) return SFB_E_BAD_ALIGN;
    if (!aligned16(cnn_out) || !aligned16(tiles) || !aligned16(workspace)) return SFB_E_BAD_ALIGN;
    if (workspace_bytes < sfb::enc_tail_workspace_bytes()) return SFB_E_WORKSPACE;
    if (tiles_bytes < sfb_enc_tail_tiles_bytes(frames, (int)N, C)) return SFB_E_WORKSPACE;
    DevInfo di;
    int rc = device_info(&di);
    if (rc) return rc;
    if (di.cc / 10 != 10) return SFB_E_UNSUPPORTED_ARCH;
    if (max_ctas > 0 && max_ctas < di.sms) di.sms = max_ctas;
    // the kernel tiles a frame in 128-pixel tiles up to ceil(N / 128); the tile buffer of a frame holds n16 / 128
    // tiles (a multiple of the pass kernel's chunk): tiles beyond ceil(N / 128) are never read with pixels < N
    sfb::SAWorkspace ws;
    sfb::sa_workspace_layout(1, 1, (int)N, C, C, 2 * C, 1, &ws);
    cudaError_t e = sfb::enc_tail_launch(cnn_out, frame_stride, frames, H, W, tiles, reinterpret_cast<const char*>(workspace),
                                         di.sms, ws.n16 / 128, (flags & SFB_ET_NHWC) != 0, reinterpret_cast<cudaStream_t>(stream));
    if (e != cudaSuccess) return cuda_err(e);
    g_launches.fetch_add(1);
    return SFB_OK;
}

// ------------------------------------------------------------------------------------------
// Rollout
// ------------------------------------------------------------------------------------------
static size_t pad128(int n) { return (size_t)((n + 127) / 128 * 128); }
// packed fp16 weights; output-feature rows of every matrix are padded to multiples of 128
static size_t ro_ws_elems(int Ds, int d, int F, int layers) {
    return 2 * (pad128(d) * Ds + pad128(Ds) * d) +        // in_proj / out_proj: hi and lo halves
           (size_t)layers * (pad128(3 * d) * d + pad128(d) * d + pad128(F) * d + pad128(d) * F);
}

// fp32 per-layer parameter blocks that follow the fp16 weights (16-byte aligned)
static size_t ro_par_floats(int d, int F) { return (size_t)9 * d + F; }
static size_t ro_par_offset(int Ds, int d, int F, int layers) {
    return (ro_ws_elems(Ds, d, F, layers) * sizeof(__half) + 255) / 256 * 256;
}

size_t sfb_rollout_workspace_bytes(int Ds, int d, int F, int num_layers) {
    if (Ds <= 0 || d <= 0 || F <= 0 || num_layers <= 0) return 0;
    return ro_par_offset(Ds, d, F, num_layers) + (size_t)num_layers * ro_par_floats(d, F) * sizeof(float);
}

// workspace layout (fp16): w_in [d][Ds] | w_out [Ds][d] | w_in_lo | w_out_lo | per layer: wqkv [3d][d], wo [d][d], w1 [F][d], w2 [d][F]
int sfb_rollout_prepare(const sfb_ro_weights* w, int Ds, int d, int F, void* workspace,
                        size_t workspace_bytes, void* stream) {
    if (!w || !workspace) return SFB_E_NULL;
    if (w->num_layers < 1 || w->num_layers > SFB_RO_MAX_LAYERS) return SFB_E_BAD_SHAPE;
    if (workspace_bytes < sfb_rollout_workspace_bytes(Ds, d, F, w->num_layers)) return SFB_E_WORKSPACE;
    if (!aligned16(workspace)) return SFB_E_BAD_ALIGN;
    if (!w->in_proj_weight || !w->out_proj_weight) return SFB_E_NULL;
    if (Ds % 64 || d % 64 || F % 64) return SFB_E_BAD_SHAPE;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    __half* dst = reinterpret_cast<__half*>(workspace);
    struct Job { const float* src; size_t n; };
    cudaError_t e;
    auto run = [&](const float* src, int N_, int K_, bool lo = false) -> int {
        const size_t n = pad128(N_) * K_;
        if (!src) return SFB_E_NULL;
        e = sfb::ro_pack2_launch(src, dst, N_, K_, st, lo);
        if (e != cudaSuccess) return cuda_err(e);
        g_launches.fetch_add(1);
        dst += n;
        return SFB_OK;
    };
    int rc;
    if ((rc = run(w->in_proj_weight, d, Ds))) return rc;
    if ((rc = run(w->out_proj_weight, Ds, d))) return rc;
    if ((rc = run(w->in_proj_weight, d, Ds, true))) return rc;
    if ((rc = run(w->out_proj_weight, Ds, d, true))) return rc;
    for (int l = 0; l < w->num_layers; ++l) {
        const sfb_ro_layer& ly = w->layers[l];
        if ((rc = run(ly.self_attn_in_proj_weight, 3 * d, d))) return rc;
        if ((rc = run(ly.self_attn_out_proj_weight, d, d))) return rc;
        if ((rc = run(ly.linear1_weight, F, d))) return rc;
        if ((rc = run(ly.linear2_weight, d, F))) return rc;
    }
    // biases + LayerNorm affine of every layer as one contiguous fp32 block (staged with one bulk copy per layer)
    float* par = reinterpret_cast<float*>(reinterpret_cast<char*>(workspace) + ro_par_offset(Ds, d, F, w->num_layers));
    for (int l = 0; l < w->num_layers; ++l) {
        const sfb_ro_layer& ly = w->layers[l];
        const float* srcs[8] = {ly.self_attn_in_proj_bias, ly.self_attn_out_proj_bias, ly.linear1_bias, ly.linear2_bias,
                                ly.norm1_weight, ly.norm1_bias, ly.norm2_weight, ly.norm2_bias};
        const int lens[8] = {3 * d, d, F, d, d, d, d, d};
        for (int i = 0; i < 8; ++i) {
            if (!srcs[i]) return SFB_E_NULL;
            e = cudaMemcpyAsync(par, srcs[i], (size_t)lens[i] * sizeof(float), cudaMemcpyDeviceToDevice, st);
            if (e != cudaSuccess) return cuda_err(e);
            par += lens[i];
        }
    }
    return SFB_OK;
}

int sfb_rollout_forward(const float* hist, float* pred_out, const sfb_ro_weights* w, int B,
                        int T_h, int K, int Ds, int d, int F, int heads, int pred_len, int mode,
                        int cond_len, unsigned int flags, const void* workspace, size_t workspace_bytes,
                        void* stream) {
    if (B == 0 || pred_len == 0) return SFB_OK;
    if (!hist || !pred_out || !w || !workspace) return SFB_E_NULL;
    if (B < 0 || T_h < 1 || K < 1 || K > 16 || pred_len < 0) return SFB_E_BAD_SHAPE;
    if (w->num_layers < 1 || w->num_layers > SFB_RO_MAX_LAYERS) return SFB_E_BAD_SHAPE;
    if (mode != SFB_RO_SLIDE && mode != SFB_RO_GROW) return SFB_E_BAD_SHAPE;
    if (mode == SFB_RO_GROW && (T_h != 1 || cond_len < 1)) return SFB_E_BAD_SHAPE;
    if (!aligned16(hist) || !aligned16(pred_out) || !aligned16(workspace)) return SFB_E_BAD_ALIGN;
    if (workspace_bytes < sfb_rollout_workspace_bytes(Ds, d, F, w->num_layers)) return SFB_E_WORKSPACE;
    if (!w->in_proj_bias || !w->out_proj_bias || !w->enc_pe) return SFB_E_NULL;
    DevInfo di;
    int rc = device_info(&di);
    if (rc) return rc;
    if (di.cc / 10 != 10) return SFB_E_UNSUPPORTED_ARCH;
    if (B == 0 || pred_len == 0) return SFB_OK;

    sfb::ROParams p{};
    p.hist = hist; p.pred = pred_out;
    const __half* ws = reinterpret_cast<const __half*>(workspace);
    p.w_in = ws; ws += pad128(d) * Ds;
    p.w_out = ws; ws += pad128(Ds) * d;
    p.w_in_lo = ws; ws += pad128(d) * Ds;
    p.w_out_lo = ws; ws += pad128(Ds) * d;
    p.b_in = w->in_proj_bias; p.b_out = w->out_proj_bias; p.pe = w->enc_pe;
    p.par_g = reinterpret_cast<const float*>(reinterpret_cast<const char*>(workspace) + ro_par_offset(Ds, d, F, w->num_layers));
    for (int l = 0; l < w->num_layers; ++l) {
        const sfb_ro_layer& s = w->layers[l];
        sfb::ROLayer& t = p.layer[l];
        t.wqkv = ws; ws += pad128(3 * d) * d;
        t.wo = ws; ws += pad128(d) * d;
        t.w1 = ws; ws += pad128(F) * d;
        t.w2 = ws; ws += pad128(d) * F;
        t.bqkv = s.self_attn_in_proj_bias; t.bo = s.self_attn_out_proj_bias;
        t.b1 = s.linear1_bias; t.b2 = s.linear2_bias;
        t.ln1w = s.norm1_weight; t.ln1b = s.norm1_bias; t.ln2w = s.norm2_weight; t.ln2b = s.norm2_bias;
        if (!t.bqkv || !t.bo || !t.b1 || !t.b2 || !t.ln1w || !t.ln1b || !t.ln2w || !t.ln2b) return SFB_E_NULL;
    }
    p.B = B; p.hist_tokens = T_h * K; p.K = K; p.Ds = Ds; p.d = d; p.F = F; p.heads = heads;
    p.layers = w->num_layers; p.pred_len = pred_len; p.mode = mode;
    p.cond_tokens = (mode == SFB_RO_GROW) ? cond_len * K : T_h * K;
    p.pe_tokens = p.cond_tokens;
    p.lmax = p.cond_tokens;
    p.prof = g_prof; p.prof_cap = g_prof_cap;
    p.dbg = debug_switches();
    size_t smem = 0;
    // engine B (tcgen05 + TMEM) when the window fits on chip, else engine A (mma.sync); SFB_RO_MMA_SYNC forces A
    const bool force_mma = (flags & SFB_RO_MMA_SYNC) != 0;
    cudaError_t e;
    if (!force_mma && sfb::ro_umma_plan(&p, di.smem_optin, &smem) == 0) {
        e = sfb::ro_umma_launch(p, smem, reinterpret_cast<cudaStream_t>(stream));
    } else {
        if (sfb::ro_mma_plan(&p, di.smem_optin, &smem)) return SFB_E_BAD_SHAPE;
        e = sfb::ro_mma_launch(p, smem, reinterpret_cast<cudaStream_t>(stream));
    }
    if (e != cudaSuccess) return cuda_err(e);
    g_launches.fetch_add(1);
    return SFB_OK;
}

// ------------------------------------------------------------------------------------------
// Decoder epilogue (section 8 f2)
// ------------------------------------------------------------------------------------------
int sfb_decode_combine(const float* dec_out, float* masks, float* recon_combined, long long* seg,
                       void* slot_max_ws, int B, int K, int HW, float fg_thre, void* stream) {
    if (B == 0) return SFB_OK;
    if (!dec_out || !masks || !recon_combined) return SFB_E_NULL;
    if (seg && !slot_max_ws) return SFB_E_NULL;
    if (B < 0 || K < 1 || K > 12 || HW < 4 || (HW & 3)) return SFB_E_BAD_SHAPE;
    if (!aligned16(dec_out) || !aligned16(masks) || !aligned16(recon_combined)) return SFB_E_BAD_ALIGN;
    DevInfo di;
    int rc = device_info(&di);
    if (rc) return rc;
    if (di.cc / 10 != 10) return SFB_E_UNSUPPORTED_ARCH;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    unsigned int* smax = seg ? reinterpret_cast<unsigned int*>(slot_max_ws) : nullptr;
    cudaError_t e;
    if (smax) {
        e = cudaMemsetAsync(smax, 0, (size_t)B * K * sizeof(unsigned int), st);
        if (e != cudaSuccess) return cuda_err(e);
    }
    // frames sit on gridDim.y (<= 65535): larger batches go in slices
    for (int b0 = 0; b0 < B; b0 += 65535) {
        const int nb = (B - b0) < 65535 ? (B - b0) : 65535;
        const size_t px = (size_t)b0 * HW;
        unsigned int* sm = smax ? smax + (size_t)b0 * K : nullptr;
        e = sfb::decode_combine_launch(dec_out + px * K * 4, masks + px * K, recon_combined + px * 3, sm, nb, K, HW,
                                       di.sms, st);
        if (e != cudaSuccess) return cuda_err(e);
        g_launches.fetch_add(1);
        if (seg) {
            e = sfb::seg_argmax_launch(masks + px * K, sm, seg + px, nb, K, HW, fg_thre, di.sms, st);
            if (e != cudaSuccess) return cuda_err(e);
            g_launches.fetch_add(1);
        }
    }
    return SFB_OK;
}

int sfb_postproc_mask(const float* masks, long long* seg, void* slot_max_ws, int B, int K, int HW, float fg_thre,
                      void* stream) {
    if (B == 0) return SFB_OK;
    if (!masks || !seg || !slot_max_ws) return SFB_E_NULL;
    if (B < 0 || K < 1 || K > 16 || HW < 1) return SFB_E_BAD_SHAPE;
    DevInfo di;
    int rc = device_info(&di);
    if (rc) return rc;
    if (di.cc / 10 != 10) return SFB_E_UNSUPPORTED_ARCH;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    unsigned int* smax = reinterpret_cast<unsigned int*>(slot_max_ws);
    cudaError_t e = cudaMemsetAsync(smax, 0, (size_t)B * K * sizeof(unsigned int), st);
    if (e != cudaSuccess) return cuda_err(e);
    // (frame, slot) planes / frames sit on gridDim.y (<= 65535): larger batches go in slices of whole frames
    const int fmax = 65535 / K;
    for (int b0 = 0; b0 < B; b0 += fmax) {
        const int nb = (B - b0) < fmax ? (B - b0) : fmax;
        const size_t px = (size_t)b0 * HW;
        if ((e = sfb::mask_max_launch(masks + px * K, smax + (size_t)b0 * K, nb * K, HW, di.sms, st)) != cudaSuccess) return cuda_err(e);
        if ((e = sfb::seg_argmax_launch(masks + px * K, smax + (size_t)b0 * K, seg + px, nb, K, HW, fg_thre, di.sms, st)) != cudaSuccess)
            return cuda_err(e);
        g_launches.fetch_add(2);
    }
    return SFB_OK;
}

}  // extern "C"

// ------------------------------------------------------------------------------------------
// SAVi slot transition (section 8 f3): program + weight layout for csrc/transition.cu
// ------------------------------------------------------------------------------------------
namespace {

struct TrLinear { long long w, b; int Kd, N; };                 // blob offsets (floats)
struct TrNorm { long long g, b; int W; };
struct TrLayout {
    TrLinear qkv[SFB_TR_MAX_LAYERS], wo[SFB_TR_MAX_LAYERS], w1[SFB_TR_MAX_LAYERS], w2[SFB_TR_MAX_LAYERS];
    TrNorm n1[SFB_TR_MAX_LAYERS], n2[SFB_TR_MAX_LAYERS];
    TrNorm mlp_ln; TrLinear mlp0, mlp2;
    TrLinear gates, outp;
    TrLinear kd0, kd3; TrNorm kdn;
    long long floats, vec_floats;
};

inline long long tr_alloc(long long* top, long long n) { const long long o = *top; *top += (n + 3) & ~3LL; return o; }

// 0 = ok; the layout depends on the structure fields and D only (prepare and forward agree on it)
int tr_layout(const sfb_tr_weights* w, int D, TrLayout* L) {
    if (!w || D < 32 || D > 512 || (D % 32)) return SFB_E_BAD_SHAPE;
    // every vector (biases, LayerNorm affine) first, contiguous: the kernel stages that region in shared memory;
    // two passes: the first one only sizes the vector region
    bool ok = true;
    long long vec_total = 0, top = 0;
    for (int pass = 0; pass < 2; ++pass) {
        long long vtop = 0;
        top = vec_total;
        auto lin = [&](TrLinear* t, int Kd, int N) {
            t->Kd = Kd; t->N = N;
            t->w = tr_alloc(&top, (long long)Kd * N);
            t->b = tr_alloc(&vtop, N);
            return (N % 32 == 0) && N <= sfb::TR_MAX_WIDTH && Kd <= sfb::TR_MAX_WIDTH && Kd >= 8;
        };
        auto norm = [&](TrNorm* t, int W) { t->W = W; t->g = tr_alloc(&vtop, W); t->b = tr_alloc(&vtop, W); };
        ok = true;
        if (w->pred_type == SFB_TR_TRANSFORMER) {
            if (w->num_layers < 1 || w->num_layers > SFB_TR_MAX_LAYERS) return SFB_E_BAD_SHAPE;
            if (w->num_heads < 1 || w->num_heads > 16 || D % w->num_heads) return SFB_E_BAD_SHAPE;
            for (int l = 0; l < w->num_layers; ++l) {
                ok &= lin(&L->qkv[l], D, 3 * D);
                ok &= lin(&L->wo[l], D, D);
                ok &= lin(&L->w1[l], D, w->ffn_dim);
                ok &= lin(&L->w2[l], w->ffn_dim, D);
                norm(&L->n1[l], D);
                norm(&L->n2[l], D);
            }
        } else if (w->pred_type == SFB_TR_MLP) {
            norm(&L->mlp_ln, D);
            ok &= lin(&L->mlp0, D, w->mlp_hidden);
            ok &= lin(&L->mlp2, w->mlp_hidden, D);
        } else if (w->pred_type != SFB_TR_NONE) {
            return SFB_E_BAD_SHAPE;
        }
        if (w->rnn_hidden > 0) {
            if (w->pred_type == SFB_TR_NONE || 2 * w->rnn_hidden > sfb::TR_MAX_WIDTH) return SFB_E_BAD_SHAPE;
            ok &= lin(&L->gates, D + w->rnn_hidden, 4 * w->rnn_hidden);
            ok &= lin(&L->outp, w->rnn_hidden, D);
        }
        ok &= lin(&L->kd0, D, 2 * D);
        if (w->kernel_mlp) {
            norm(&L->kdn, 2 * D);
            ok &= lin(&L->kd3, 2 * D, 2 * D);
        }
        vec_total = vtop;
    }
    if (vec_total > sfb::TR_MAX_VEC) return SFB_E_BAD_SHAPE;
    L->vec_floats = vec_total;
    L->floats = top;
    return ok ? SFB_OK : SFB_E_BAD_SHAPE;
}

// The program of one transition.  LINEAR steps end in a cluster barrier and write their result into the buffer of every
// CTA of the cluster: their destination must not have been read or written by any step since the previous barrier
// (a slower CTA may still be there).  `touched` tracks that; a violation is a bug in this builder, reported as BAD_SHAPE.
struct TrBuilder {
    sfb::TrParams* p;
    unsigned touched = 0, reserved = 0;
    int err = 0;
    void push(int code, int a0 = 0, int a1 = 0, int a2 = 0, int a3 = 0, int a4 = 0, int a5 = 0, int a6 = 0) {
        if (p->nops >= sfb::TR_MAX_OPS) { err = 1; return; }
        sfb::TrOp& o = p->ops[p->nops++];
        o.code = code; o.a[0] = a0; o.a[1] = a1; o.a[2] = a2; o.a[3] = a3; o.a[4] = a4; o.a[5] = a5; o.a[6] = a6;
    }
    int pick(unsigned live) {
        for (int i = 0; i < sfb::TR_NBUF; ++i) if (!(((touched | live | reserved) >> i) & 1u)) return i;
        err = 1; return 0;
    }
    void load(int buf, int in, int W, int rstride, int foff, long long cstride, int flags) {
        push(sfb::TR_LOAD, buf, in, W, rstride, foff, (int)cstride, flags); touched |= 1u << buf;
    }
    void ln(int src, int dst, const TrNorm& n, int flags) {
        push(sfb::TR_LN, src, dst, n.W, (int)n.g, (int)n.b, flags); touched |= (1u << src) | (1u << dst);
    }
    // src2 >= 0: operand rows k >= ksplit come from buffer src2 at feature offset off2
    void linear(int src, int dst, const TrLinear& t, int flags, int src2 = -1, int off2 = 0, int ksplit = 0) {
        if (src == dst || src2 == dst || ((touched >> dst) & 1u)) err = 1;
        int kd = t.Kd;
        if (src2 >= 0) { kd |= ksplit << 16; flags |= (src2 << 8) | (off2 << 12); }
        push(sfb::TR_LINEAR, src, dst, kd, t.N, (int)t.w, (int)t.b, flags);
        touched = 0;                     // cluster barrier
    }
    void attn(int src, int dst, int D, int heads) { push(sfb::TR_ATTN, src, dst, D, heads); touched |= (1u << src) | (1u << dst); }
    void lstm(int g, int c, int h, int H) { push(sfb::TR_LSTM, g, c, h, H); touched |= (1u << g) | (1u << c) | (1u << h); }
};

int tr_program(const sfb_tr_weights* w, int D, const TrLayout& L, int K, long long prev_clip_stride, int use_predictor,
               sfb::TrParams* p) {
    TrBuilder b{p};
    p->nops = 0;
    if (L.floats >= (1LL << 31)) return SFB_E_BAD_SHAPE;
    if (prev_clip_stride >= (1LL << 31) || prev_clip_stride < 0) return SFB_E_BAD_SHAPE;
    const bool rnn = use_predictor && w->rnn_hidden > 0;
    const int H = w->rnn_hidden, SB = sfb::TR_NBUF - 1;       // the LSTM state waits in the last buffer: [c | h]
    int cur = 0;
    // every input of the step is fetched up front (one exposed global-load latency instead of three)
    b.load(cur, 0, D, D, 0, prev_clip_stride, rnn ? sfb::TR_F_NOSYNC : 0);
    if (rnn) {
        b.load(SB, 1, H, H, H, -1, sfb::TR_F_NOSYNC);
        b.load(SB, 2, H, H, 0, -1, 0);
        b.reserved = 1u << SB;
    }
    auto bit = [](int i) { return 1u << i; };
    if (use_predictor && w->pred_type == SFB_TR_TRANSFORMER) {
        for (int l = 0; l < w->num_layers; ++l) {
            if (w->norm_first) {
                int y = b.pick(bit(cur));
                b.ln(cur, y, L.n1[l], 0);
                int z = b.pick(bit(cur) | bit(y));
                b.linear(y, z, L.qkv[l], 0);
                y = b.pick(bit(cur) | bit(z));
                b.attn(z, y, D, w->num_heads);
                b.linear(y, cur, L.wo[l], sfb::TR_F_ADD);
                y = b.pick(bit(cur));
                b.ln(cur, y, L.n2[l], 0);
                z = b.pick(bit(cur) | bit(y));
                b.linear(y, z, L.w1[l], sfb::TR_F_RELU);
                b.linear(z, cur, L.w2[l], sfb::TR_F_ADD);
            } else {
                int z = b.pick(bit(cur));
                b.linear(cur, z, L.qkv[l], 0);
                int y = b.pick(bit(cur) | bit(z));
                b.attn(z, y, D, w->num_heads);
                b.linear(y, cur, L.wo[l], sfb::TR_F_ADD);
                b.ln(cur, cur, L.n1[l], 0);
                z = b.pick(bit(cur));
                b.linear(cur, z, L.w1[l], sfb::TR_F_RELU);
                b.linear(z, cur, L.w2[l], sfb::TR_F_ADD);
                b.ln(cur, cur, L.n2[l], 0);
            }
        }
    } else if (use_predictor && w->pred_type == SFB_TR_MLP) {
        int y = b.pick(bit(cur));
        b.ln(cur, y, L.mlp_ln, 0);
        int z = b.pick(bit(cur) | bit(y));
        b.linear(y, z, L.mlp0, sfb::TR_F_RELU);
        const int skip = w->norm_first ? y : cur;
        b.linear(z, skip, L.mlp2, sfb::TR_F_ADD);
        cur = skip;
    }
    if (rnn) {
        int z = b.pick(bit(cur));
        b.linear(cur, z, L.gates, 0, SB, H, D);            // gates = [x ; h] [W_ih | W_hh]^T + b_ih + b_hh
        int hb = b.pick(bit(z));
        b.lstm(z, SB, hb, H);
        b.reserved = 0;
        int y = b.pick(bit(hb));
        b.linear(hb, y, L.outp, 0);
        cur = y;
    }
    {
        int z = b.pick(bit(cur));
        b.linear(cur, z, L.kd0, 0);
        cur = z;
        if (w->kernel_mlp) {
            b.ln(cur, cur, L.kdn, sfb::TR_F_RELU);
            int x = b.pick(bit(cur));
            b.linear(cur, x, L.kd3, 0);
            cur = x;
        }
    }
    b.push(sfb::TR_STORE, cur, 0, 2 * D, 2 * D, 0, 0, sfb::TR_F_NOSYNC);
    b.push(sfb::TR_SAMPLE, cur, D);
    (void)K;
    return b.err ? SFB_E_BAD_SHAPE : SFB_OK;
}

}  // namespace

extern "C" {

size_t sfb_transition_workspace_bytes(const sfb_tr_weights* w, int D) {
    TrLayout L;
    if (tr_layout(w, D, &L)) return 0;
    return (size_t)L.floats * sizeof(float);
}

int sfb_transition_prepare(const sfb_tr_weights* w, int D, void* workspace, size_t workspace_bytes, void* stream) {
    TrLayout L;
    int rc = tr_layout(w, D, &L);
    if (rc) return rc;
    if (!workspace) return SFB_E_NULL;
    if (workspace_bytes < (size_t)L.floats * sizeof(float)) return SFB_E_WORKSPACE;
    if (!aligned16(workspace)) return SFB_E_BAD_ALIGN;
    float* blob = reinterpret_cast<float*>(workspace);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    cudaError_t e = cudaSuccess;
    int missing = 0;
    auto lin = [&](const TrLinear& t, const float* W, const float* bias) {
        if (!W || !bias) { missing = 1; return; }
        if (e == cudaSuccess) e = sfb::tr_pack_weight(W, blob + t.w, t.N, t.Kd, 0, st);
        if (e == cudaSuccess) e = sfb::tr_pack_vec(bias, nullptr, blob + t.b, t.N, st);
    };
    auto norm = [&](const TrNorm& t, const float* g, const float* b) {
        if (!g || !b) { missing = 1; return; }
        if (e == cudaSuccess) e = sfb::tr_pack_vec(g, nullptr, blob + t.g, t.W, st);
        if (e == cudaSuccess) e = sfb::tr_pack_vec(b, nullptr, blob + t.b, t.W, st);
    };
    if (w->pred_type == SFB_TR_TRANSFORMER) {
        for (int l = 0; l < w->num_layers; ++l) {
            const sfb_ro_layer& s = w->layers[l];
            lin(L.qkv[l], s.self_attn_in_proj_weight, s.self_attn_in_proj_bias);
            lin(L.wo[l], s.self_attn_out_proj_weight, s.self_attn_out_proj_bias);
            lin(L.w1[l], s.linear1_weight, s.linear1_bias);
            lin(L.w2[l], s.linear2_weight, s.linear2_bias);
            norm(L.n1[l], s.norm1_weight, s.norm1_bias);
            norm(L.n2[l], s.norm2_weight, s.norm2_bias);
        }
    } else if (w->pred_type == SFB_TR_MLP) {
        norm(L.mlp_ln, w->ln_weight, w->ln_bias);
        lin(L.mlp0, w->mlp_0_weight, w->mlp_0_bias);
        lin(L.mlp2, w->mlp_2_weight, w->mlp_2_bias);
    }
    if (w->rnn_hidden > 0) {
        const int H = w->rnn_hidden;
        if (!w->rnn_weight_ih_l0 || !w->rnn_weight_hh_l0 || !w->rnn_bias_ih_l0 || !w->rnn_bias_hh_l0) missing = 1;
        else {
            // gates = [x ; h] [W_ih | W_hh]^T + (b_ih + b_hh): one k-major matrix of D + H rows
            if (e == cudaSuccess) e = sfb::tr_pack_weight(w->rnn_weight_ih_l0, blob + L.gates.w, 4 * H, D, 0, st);
            if (e == cudaSuccess) e = sfb::tr_pack_weight(w->rnn_weight_hh_l0, blob + L.gates.w, 4 * H, H, D, st);
            if (e == cudaSuccess) e = sfb::tr_pack_vec(w->rnn_bias_ih_l0, w->rnn_bias_hh_l0, blob + L.gates.b, 4 * H, st);
        }
        lin(L.outp, w->out_projector_weight, w->out_projector_bias);
    }
    lin(L.kd0, w->kernel_dist_0_weight, w->kernel_dist_0_bias);
    if (w->kernel_mlp) {
        norm(L.kdn, w->kernel_dist_1_weight, w->kernel_dist_1_bias);
        lin(L.kd3, w->kernel_dist_3_weight, w->kernel_dist_3_bias);
    }
    if (missing) return SFB_E_NULL;
    if (e != cudaSuccess) return cuda_err(e);
    return SFB_OK;
}

int sfb_transition_forward(const sfb_tr_weights* w, int D, int B, int K, const float* prev, long long prev_clip_stride,
                           int use_predictor, const float* h_in, const float* c_in, const float* noise,
                           float* dist_out, float* slots_out, float* h_out, float* c_out, const void* workspace,
                           size_t workspace_bytes, void* stream) {
    if (B == 0) return SFB_OK;
    TrLayout L;
    int rc = tr_layout(w, D, &L);
    if (rc) return rc;
    if (!prev || !workspace || !dist_out || !slots_out) return SFB_E_NULL;
    if (B < 0 || K < 1 || K > 8) return SFB_E_BAD_SHAPE;
    if (workspace_bytes < (size_t)L.floats * sizeof(float)) return SFB_E_WORKSPACE;
    if (!aligned16(workspace)) return SFB_E_BAD_ALIGN;
    if (use_predictor && w->rnn_hidden > 0 && (!h_out || !c_out)) return SFB_E_NULL;
    DevInfo di;
    rc = device_info(&di);
    if (rc) return rc;
    if (di.cc / 10 != 10) return SFB_E_UNSUPPORTED_ARCH;
    sfb::TrParams p{};
    p.blob = reinterpret_cast<const float*>(workspace);
    p.in[0] = prev; p.in[1] = h_in; p.in[2] = c_in; p.in[3] = noise;
    p.out[0] = dist_out; p.out[1] = slots_out; p.out[2] = h_out; p.out[3] = c_out;
    p.B = B; p.K = K;
    p.prof = g_prof; p.prof_cap = g_prof_cap;
    p.vec_floats = (int)L.vec_floats;
    rc = tr_program(w, D, L, K, prev_clip_stride, use_predictor, &p);
    if (rc) return rc;
    // one cluster per clip; the cluster splits every LINEAR by output feature (all widths are multiples of 32).  The
    // largest cluster size whose clusters are all resident at once (a second wave would double the latency).
    static std::atomic<int> max_clusters[64][3][4];          // [device][row variant][log2 cluster size], 0 = not asked yet
    int dev = 0;
    cudaGetDevice(&dev);
    const int rv = K <= 4 ? 0 : (K <= 6 ? 1 : 2);
    int csize = 1;
    for (int lg = 3; lg >= 1; --lg) {
        int n = (dev >= 0 && dev < 64) ? max_clusters[dev][rv][lg].load() : 0;
        if (n == 0) {
            n = sfb::transition_max_clusters(K, 1 << lg);
            if (n <= 0) n = -1;
            if (dev >= 0 && dev < 64) max_clusters[dev][rv][lg].store(n);
        }
        if (n >= B) { csize = 1 << lg; break; }
    }
    cudaError_t e = sfb::transition_launch(p, csize, reinterpret_cast<cudaStream_t>(stream));
    if (e != cudaSuccess) return cuda_err(e);
    g_launches.fetch_add(1);
    return SFB_OK;
}

}  // extern "C"
