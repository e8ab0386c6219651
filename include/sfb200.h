/* sfb200.h -- C ABI of libsfb200.so, the sm_100a engine behind the SlotFormer
 * hot paths.
 *
 * The reference (pairlab/SlotFormer) is pure Python/PyTorch and has no FFI of
 * its own; this header is the thin extension boundary
 * BASELINE.json's north_star prescribes.  Each entry point names the reference
 * operator it replaces (paths relative to /root/reference/slotformer):
 *
 *   sfb_sa_forward        SlotAttention.forward        base_slots/models/savi.py:56-102
 *                         SlotAttentionWMask.forward   base_slots/models/steve.py:19-73
 *   sfb_rollout_forward   SlotRollouter.forward        video_prediction/models/slotformer.py:85-126
 *                         SingleStepSlotRollouter.forward
 *                                                      video_prediction/models/single_step_slotformer.py:49-90
 *
 * Conventions
 *   - every function returns int: 0 = ok, <0 = error (see SFB_E_*); nothing throws or aborts;
 *   - every data pointer is a DEVICE pointer borrowed for the duration of the (asynchronous)
 *     call; the library allocates nothing persistent -- scratch is caller-provided workspace;
 *   - `stream` is a cudaStream_t passed as void*; calls only enqueue work on it, never
 *     synchronise the host (except the *_host convenience entry points, which say so);
 *   - weight pointers are the reference state_dict tensors themselves (fp32, contiguous,
 *     PyTorch [out_features, in_features] layout), field names = state_dict keys.
 */
#ifndef SFB200_H_
#define SFB200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SFB_VERSION 210 /* 0.2.1: sfb_transition_*, flags argument of sfb_enc_tail_forward */

#define SFB_OK 0
#define SFB_E_BAD_SHAPE (-1)        /* unsupported / inconsistent dimensions            */
#define SFB_E_BAD_ALIGN (-2)        /* pointer or stride not 16-byte aligned            */
#define SFB_E_UNSUPPORTED_ARCH (-3) /* device is not sm_100                             */
#define SFB_E_WORKSPACE (-4)        /* workspace too small                              */
#define SFB_E_NULL (-5)             /* required pointer is NULL                         */
#define SFB_E_CUDA_BASE (-100)      /* -100 - cudaError_t                               */

#define SFB_DTYPE_F32 0
#define SFB_DTYPE_BF16 1
#define SFB_DTYPE_TILES16 2 /* pre-normalised fp16 operand tiles written by sfb_enc_tail_forward (C = 128) */

#define SFB_RO_SLIDE 0 /* SlotRollouter: fixed window, drop oldest frame each step       */
#define SFB_RO_GROW 1  /* SingleStepSlotRollouter: window grows up to cond_len frames    */

int sfb_version(void);
const char* sfb_strerror(int code);
/* number of kernel launches this process has made through the library (bench.py's gpu_launches) */
long long sfb_launch_count(void);

/* The library keeps no mutable state between calls (sfb_launch_count is a statistic nobody reads back): every
 * entry point is re-entrant across host threads, streams and devices, and safe to capture into CUDA graphs. */

#ifdef SFB_DEBUG
/* Debug / profiling aids: only in the debug build (python -m slotformer_b200.build --debug ->
 * libsfb200_debug.so), never in the product library.
 * sfb_debug_set_profile: device buffer of `capacity` uint64 that the NEXT forward calls fill with
 * %globaltimer stamps at phase boundaries (cluster 0 / CTA 0); pass NULL to switch it off.  The debug build also
 * honours the SFB_DBG environment switches that disable parts of a kernel for timing experiments.
 */
void sfb_debug_set_profile(void* device_buf, int capacity);
/* self-test of the tcgen05 path: out[N][M] = X[N][K] W[M][K]^T (fp16 operands, fp32 accumulate);
 * M % 128 == 0, 16 <= N <= 128 (N % 16 == 0), K % 64 == 0; workspace >= M*K*2 bytes. */
int sfb_debug_umma_gemm(const float* W, const float* X, float* out, int M, int N, int K, void* workspace,
                        size_t workspace_bytes, void* stream);
#endif

/* ------------------------------------------------------------------------- */
/* Hot path 1: Slot Attention                                                 */
/* ------------------------------------------------------------------------- */
/* Parameter set of reference SlotAttention.__init__ (savi.py:19-54). */
typedef struct sfb_sa_weights {
    const float* norm_inputs_weight; /* [C]    */
    const float* norm_inputs_bias;   /* [C]    */
    const float* project_q_0_weight; /* [D]   LayerNorm */
    const float* project_q_0_bias;   /* [D]    */
    const float* project_q_1_weight; /* [D,D] no bias   */
    const float* project_k_weight;   /* [D,C] no bias   */
    const float* project_v_weight;   /* [D,C] no bias   */
    const float* gru_weight_ih;      /* [3D,D] gates r,z,n */
    const float* gru_weight_hh;      /* [3D,D] */
    const float* gru_bias_ih;        /* [3D]   */
    const float* gru_bias_hh;        /* [3D]   */
    const float* mlp_0_weight;       /* [D]   LayerNorm */
    const float* mlp_0_bias;         /* [D]    */
    const float* mlp_1_weight;       /* [Dm,D] */
    const float* mlp_1_bias;         /* [Dm]   */
    const float* mlp_3_weight;       /* [D,Dm] */
    const float* mlp_3_bias;         /* [D]    */
} sfb_sa_weights;

/* Bytes of device workspace sfb_sa_forward needs: folded fp16 hi/lo weights, per-frame q~,
 * per-(frame, pixel-chunk) partial sums and the fp16 x^ ring of one frame chunk. */
size_t sfb_sa_workspace_bytes(int B, int N, int C, int D, int Dm, int n_iter, int chunk_frames);

/* Fold the projections (W_qk = scale*log2e*gamma*(Wq^T Wk)^T, W_iv = W_ih Wv, logit-bias vector) and
 * pack all slot-update weights as swizzled fp16 hi/lo panels into the head of `workspace`.  Call
 * again whenever a weight tensor changed or the workspace moved; sfb_sa_forward only reads it. */
int sfb_sa_prepare(const sfb_sa_weights* w, int C, int D, int Dm, void* workspace, size_t workspace_bytes,
                   void* stream);

/* sfb_sa_forward flags */
#define SFB_SA_NO_TCGEN05 1u /* run the mma.sync passes (sa_pass.cu) instead of the tcgen05 ones (sa_pass_tc.cu)  */
#define SFB_SA_SPLIT_ON 2u   /* mma.sync first pass on warp pairs: force on ...                                   */
#define SFB_SA_SPLIT_OFF 4u  /* ... / off (default: on when max_ctas caps the grid)                                */
#define SFB_SA_XHAT_KEEP 8u  /* x^ ring stores keep their evict_last L2 policy under a capped grid (default: evict_first
                                when the ring exceeds L2 and max_ctas caps the grid, i.e. the GPU is shared)          */

/* Slot Attention forward for B independent frames (after sfb_sa_prepare on the same workspace).
 *   feats        [B, N, C]  fp32 (SFB_DTYPE_F32) or bf16 (SFB_DTYPE_BF16); rows contiguous, frame b at
 *                feats + b*feat_batch_stride elements (savi.py:406 passes encoder_out[:, idx]);
 *                or, with SFB_DTYPE_TILES16, the tiles of sfb_enc_tail_forward (frame b at feats +
 *                b*feat_batch_stride fp16 elements): the features are then already LayerNorm-ed operand
 *                tiles and every iteration streams them straight into the tensor cores
 *   slots_in     [B, K, D]  fp32 initial slots          slots_out [B, K, D] fp32
 *   seg_mask     NULL, or [B, K, N] fp32: softmax-over-slots attention of the LAST iteration,
 *                before +eps / renormalisation (steve.py:54-55)
 *   workspace    >= sfb_sa_workspace_bytes(...) for the same arguments, 16-byte aligned
 *   chunk_frames 0 = choose; otherwise frames processed per scheduling chunk (the fp16 x^ ring
 *                of one chunk is what later iterations re-read; keep it inside L2)
 *   max_ctas     0 = one persistent CTA per SM; otherwise the grid of the streaming passes is capped at this
 *                many CTAs, e.g. while the rollout of the previous clip batch holds one SM per clip on another
 *                stream (slotformer_b200.engine.HotPathPipeline).  Per call: nothing process-wide.
 *   flags        SFB_SA_* (0 = defaults)
 * Supported: (C, D, Dm) in {(128,128,256), (192,192,384)}, 1 <= K <= 8, n_iter >= 1, N >= 1.
 */
int sfb_sa_forward(const void* feats, int feat_dtype, int64_t feat_batch_stride,
                   const float* slots_in, float* slots_out, float* seg_mask,
                   const sfb_sa_weights* w, int B, int N, int C, int D, int Dm, int K,
                   int n_iter, float eps, int chunk_frames, int max_ctas, unsigned int flags,
                   void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------- */
/* Encoder tail (next row f1): CNN output -> Slot Attention operand tiles     */
/* ------------------------------------------------------------------------- */
/* StoSAVi._get_encoder_out after the CNN (base_slots/models/savi.py:371-377, utils.py:52-63) fused with
 * SlotAttention.norm_inputs (savi.py:66):  + SoftPositionEmbed -> LayerNorm(64) -> Linear(64, C) -> ReLU ->
 * Linear(C, C) -> LayerNorm statistics of Slot Attention, written as fp16 operand tiles.  The fp32
 * [frames, N, C] feature grid of the reference is never materialised.  C = 128, 64 CNN channels. */
typedef struct sfb_enc_tail_weights {
    const float* encoder_pos_embedding_dense_weight; /* [64, 4]  */
    const float* encoder_pos_embedding_dense_bias;   /* [64]     */
    const float* encoder_out_layer_0_weight;         /* [64]  LayerNorm */
    const float* encoder_out_layer_0_bias;           /* [64]     */
    const float* encoder_out_layer_1_weight;         /* [C, 64]  */
    const float* encoder_out_layer_1_bias;           /* [C]      */
    const float* encoder_out_layer_3_weight;         /* [C, C]   */
    const float* encoder_out_layer_3_bias;           /* [C]      */
} sfb_enc_tail_weights;

size_t sfb_enc_tail_workspace_bytes(int C);
/* bytes of the tile buffer for `frames` frames of N pixels (128-pixel tiles of 32 KB; = frames * tile_frame_bytes) */
size_t sfb_enc_tail_tiles_bytes(int frames, int N, int C);
/* fold + pack the weights into `workspace` (call again when a weight changed) */
int sfb_enc_tail_prepare(const sfb_enc_tail_weights* w, int C, void* workspace, size_t workspace_bytes, void* stream);
/* sfb_enc_tail_forward flags */
#define SFB_ET_NHWC 1u /* cnn_out is channels-last: [frames, H, W, 64] in memory (what cuDNN's tensor-core convolutions
                          write natively, torch.channels_last) instead of [frames, 64, H, W] */
/* cnn_out [frames, 64, H, W] fp32 (NCHW memory, or NHWC with SFB_ET_NHWC; frame f at cnn_out + f*frame_stride elements)
 * -> tiles (frame-major, sfb_enc_tail_tiles_bytes(1, H*W, C) bytes per frame), to be passed to sfb_sa_forward as
 * SFB_DTYPE_TILES16.  max_ctas: as in sfb_sa_forward. */
int sfb_enc_tail_forward(const float* cnn_out, int64_t frame_stride, int frames, int H, int W, int C, void* tiles,
                         size_t tiles_bytes, const void* workspace, size_t workspace_bytes, int max_ctas,
                         unsigned int flags, void* stream);

/* ------------------------------------------------------------------------- */
/* Hot path 2: autoregressive slot-Transformer rollout                        */
/* ------------------------------------------------------------------------- */
/* One nn.TransformerEncoderLayer(norm_first=True) (slotformer.py:72-78). */
typedef struct sfb_ro_layer {
    const float* self_attn_in_proj_weight;  /* [3d, d] packed q,k,v */
    const float* self_attn_in_proj_bias;    /* [3d]    */
    const float* self_attn_out_proj_weight; /* [d, d]  */
    const float* self_attn_out_proj_bias;   /* [d]     */
    const float* linear1_weight;            /* [F, d]  */
    const float* linear1_bias;              /* [F]     */
    const float* linear2_weight;            /* [d, F]  */
    const float* linear2_bias;              /* [d]     */
    const float* norm1_weight;              /* [d]     */
    const float* norm1_bias;                /* [d]     */
    const float* norm2_weight;              /* [d]     */
    const float* norm2_bias;                /* [d]     */
} sfb_ro_layer;

#define SFB_RO_MAX_LAYERS 16

/* Parameter set of reference SlotRollouter.__init__ (slotformer.py:51-83). */
typedef struct sfb_ro_weights {
    const float* in_proj_weight;  /* [d, Ds] */
    const float* in_proj_bias;    /* [d]     */
    const float* out_proj_weight; /* [Ds, d] */
    const float* out_proj_bias;   /* [Ds]    */
    /* per-token positional table [pe_frames*K, d]: enc_t_pe repeated per slot (+ enc_slots_pe
     * when configured), exactly the `enc_pe` tensor built at slotformer.py:103-110 */
    const float* enc_pe;
    int num_layers;
    sfb_ro_layer layers[SFB_RO_MAX_LAYERS];
} sfb_ro_weights;

/* Bytes of device workspace for the fp16 operand copies of the weights. */
size_t sfb_rollout_workspace_bytes(int Ds, int d, int F, int num_layers);

/* Convert the fp32 weights into the fp16 operand copies held in `workspace`.  Call again
 * whenever a weight tensor changed; sfb_rollout_forward only reads the workspace + biases. */
int sfb_rollout_prepare(const sfb_ro_weights* w, int Ds, int d, int F, void* workspace,
                        size_t workspace_bytes, void* stream);

/* sfb_rollout_forward flags */
#define SFB_RO_MMA_SYNC 1u /* force engine A (mma.sync) even when the window fits the tcgen05 engine */

/* Autoregressive rollout for B independent clips.
 *   hist      [B, T_h, K, Ds] fp32 burn-in slots        pred_out [B, pred_len, K, Ds] fp32
 *   mode      SFB_RO_SLIDE: window = T_h frames (pe_frames = T_h)
 *             SFB_RO_GROW : T_h must be 1, window grows to cond_len frames (pe_frames = cond_len),
 *                           positional rows are the LAST rows of the table
 *   Supported: Ds in {128,192}, d in {128,256}, d/heads in {16,32}, F % 64 == 0,
 *              window tokens <= 128.
 */
int sfb_rollout_forward(const float* hist, float* pred_out, const sfb_ro_weights* w, int B,
                        int T_h, int K, int Ds, int d, int F, int heads, int pred_len, int mode,
                        int cond_len, unsigned int flags, const void* workspace, size_t workspace_bytes,
                        void* stream);

/* ------------------------------------------------------------------------- */
/* SAVi slot transition (next row f3): predictor + kernel_dist_layer + sample   */
/* ------------------------------------------------------------------------- */
/* The per-frame glue between two Slot Attention calls of StoSAVi.encode (base_slots/models/savi.py:393-410):
 *   latents = predictor(prev_slots)        predictor.py:20-44 TransformerPredictor | :47-74 ResidualMLPPredictor,
 *                                          optionally inside predictor.py:76-113 RNNPredictorWrapper (LSTM, 1 layer)
 *   dist    = kernel_dist_layer(latents)   savi.py:200-212
 *   slots0  = mu + noise * exp(logvar/2)   savi.py:355-363 (noise == NULL: slots0 = mu, kld_method 'none')
 * as ONE kernel launch (a thread-block cluster per clip; csrc/transition.cu).  Field names = state_dict keys of
 * `predictor.` / `kernel_dist_layer.`; unused pointers stay NULL. */
#define SFB_TR_NONE 0        /* no predictor: latents = input rows (first frame: init_latents, prev_clip_stride 0) */
#define SFB_TR_TRANSFORMER 1 /* [base_predictor.]transformer_encoder.layers.<i>.* */
#define SFB_TR_MLP 2         /* [base_predictor.]ln.*, mlp.0.*, mlp.2.*  (channels [D, mlp_hidden, D]) */
#define SFB_TR_MAX_LAYERS 4

typedef struct sfb_tr_weights {
    int pred_type;  /* SFB_TR_* */
    int num_layers; /* transformer layers */
    int num_heads;
    int ffn_dim;
    int norm_first;
    sfb_ro_layer layers[SFB_TR_MAX_LAYERS];
    int mlp_hidden;
    const float* ln_weight;    /* [D] */
    const float* ln_bias;      /* [D] */
    const float* mlp_0_weight; /* [mlp_hidden, D] */
    const float* mlp_0_bias;
    const float* mlp_2_weight; /* [D, mlp_hidden] */
    const float* mlp_2_bias;
    int rnn_hidden;                     /* 0: no RNNPredictorWrapper */
    const float* rnn_weight_ih_l0;      /* [4H, D]  gate order i, f, g, o */
    const float* rnn_weight_hh_l0;      /* [4H, H]  */
    const float* rnn_bias_ih_l0;        /* [4H]     */
    const float* rnn_bias_hh_l0;        /* [4H]     */
    const float* out_projector_weight;  /* [D, H]   */
    const float* out_projector_bias;    /* [D]      */
    int kernel_mlp;                     /* 1: Linear -> LayerNorm -> ReLU -> Linear;  0: Linear */
    const float* kernel_dist_0_weight;  /* [2D, D]  */
    const float* kernel_dist_0_bias;
    const float* kernel_dist_1_weight;  /* LayerNorm [2D] (kernel_mlp) */
    const float* kernel_dist_1_bias;
    const float* kernel_dist_3_weight;  /* [2D, 2D] (kernel_mlp) */
    const float* kernel_dist_3_bias;
} sfb_tr_weights;

/* Bytes of device workspace for the k-major fp32 copies of the weights (0: unsupported structure / sizes). */
size_t sfb_transition_workspace_bytes(const sfb_tr_weights* w, int D);
/* Re-pack the weights into `workspace`; call again whenever a weight tensor changed. */
int sfb_transition_prepare(const sfb_tr_weights* w, int D, void* workspace, size_t workspace_bytes, void* stream);
/* One transition for B clips of K <= 8 slots (D = slot size):
 *   prev [.., K, D] fp32 rows of clip b at prev + b * prev_clip_stride (floats; 0 = the same K rows for every clip)
 *   use_predictor 0: latents = prev (first frame);  1: latents = predictor(prev)
 *   h_in / c_in  [B*K, H] LSTM state before the step (NULL = zeros), h_out / c_out [B*K, H] after it (rnn_hidden > 0)
 *   noise        NULL or [B, K, D]      dist_out [B, K, 2D]      slots_out [B, K, D]
 * Supported: D % 4 == 0, every layer width (3D, ffn_dim, mlp_hidden, 4H, D + H, 2D) <= 1024 and a multiple of 32. */
int sfb_transition_forward(const sfb_tr_weights* w, int D, int B, int K, const float* prev, long long prev_clip_stride,
                           int use_predictor, const float* h_in, const float* c_in, const float* noise,
                           float* dist_out, float* slots_out, float* h_out, float* c_out, const void* workspace,
                           size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------- */
/* Decoder epilogue (next row f2): the tail of StoSAVi.decode + postproc_mask   */
/* ------------------------------------------------------------------------- */
/* Everything after the deconvolution stack of the spatial-broadcast decoder (base_slots/models/savi.py:519-523):
 *   dec_out        [B, K, 4, HW] fp32   decoder output, planes 0..2 = colour, plane 3 = mask logit
 *   masks          [B, K, HW]    fp32   softmax over the K slots of plane 3
 *   recon_combined [B, 3, HW]    fp32   sum_k colour_k * masks_k
 *   seg            NULL, or [B, HW] int64: postproc_mask (video_prediction/vp_utils.py:20-41) of `masks` --
 *                  the frame's background slot (smallest per-slot maximum) takes every pixel whose best score is
 *                  below fg_thre, then argmax over slots (first maximum, as torch.argmax)
 *   slot_max_ws    B*K uint32 of scratch, required iff seg != NULL
 *                  (scores of any sign: the per-slot maxima use an order-preserving integer encoding)
 * Supported: 1 <= K <= 12, HW % 4 == 0; any B (batches beyond 65535 frames are sliced internally). */
int sfb_decode_combine(const float* dec_out, float* masks, float* recon_combined, long long* seg,
                       void* slot_max_ws, int B, int K, int HW, float fg_thre, void* stream);

/* postproc_mask on given masks [B, K, HW] (vp_utils.py:20-41) -> seg [B, HW] int64.  K <= 16, any B. */
int sfb_postproc_mask(const float* masks, long long* seg, void* slot_max_ws, int B, int K, int HW, float fg_thre,
                      void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SFB200_H_ */
