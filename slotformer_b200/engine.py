"""ctypes binding of libsfb200.so (include/sfb200.h) for torch tensors.

PyTorch is used for device memory and streams only; every hot-path computation is a call
into the hand-written sm_100a kernels.  There is no CPU / eager fallback: if the library is
missing or the tensors are not on a CUDA device the functions raise.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, 'lib', 'libsfb200.so')
_LIB_DEBUG_PATH = os.path.join(_HERE, 'lib', 'libsfb200_debug.so')
_lib = None
_lib_debug = None

# sfb_sa_forward / sfb_rollout_forward flags (include/sfb200.h)
SFB_SA_NO_TCGEN05 = 1
SFB_SA_SPLIT_ON = 2
SFB_SA_SPLIT_OFF = 4
SFB_SA_XHAT_KEEP = 8
SFB_RO_MMA_SYNC = 1
SFB_ET_NHWC = 1

SFB_DTYPE_F32 = 0
SFB_DTYPE_BF16 = 1
SFB_DTYPE_TILES16 = 2
SFB_RO_SLIDE = 0
SFB_RO_GROW = 1
RO_MAX_LAYERS = 16

SA_WEIGHT_KEYS = (
    'norm_inputs.weight', 'norm_inputs.bias', 'project_q.0.weight', 'project_q.0.bias',
    'project_q.1.weight', 'project_k.weight', 'project_v.weight', 'gru.weight_ih',
    'gru.weight_hh', 'gru.bias_ih', 'gru.bias_hh', 'mlp.0.weight', 'mlp.0.bias',
    'mlp.1.weight', 'mlp.1.bias', 'mlp.3.weight', 'mlp.3.bias')

ENC_TAIL_KEYS = (
    'encoder_pos_embedding.dense.weight', 'encoder_pos_embedding.dense.bias',
    'encoder_out_layer.0.weight', 'encoder_out_layer.0.bias', 'encoder_out_layer.1.weight',
    'encoder_out_layer.1.bias', 'encoder_out_layer.3.weight', 'encoder_out_layer.3.bias')

RO_LAYER_KEYS = (
    'self_attn.in_proj_weight', 'self_attn.in_proj_bias', 'self_attn.out_proj.weight',
    'self_attn.out_proj.bias', 'linear1.weight', 'linear1.bias', 'linear2.weight',
    'linear2.bias', 'norm1.weight', 'norm1.bias', 'norm2.weight', 'norm2.bias')


class SfbError(RuntimeError):
    pass


class _SAWeights(ctypes.Structure):
    _fields_ = [(k.replace('.', '_'), ctypes.c_void_p) for k in SA_WEIGHT_KEYS]


class _EncTailWeights(ctypes.Structure):
    _fields_ = [(k.replace('.', '_'), ctypes.c_void_p) for k in ENC_TAIL_KEYS]


class _ROLayer(ctypes.Structure):
    _fields_ = [(k.replace('.', '_'), ctypes.c_void_p) for k in RO_LAYER_KEYS]


class _ROWeights(ctypes.Structure):
    _fields_ = [('in_proj_weight', ctypes.c_void_p), ('in_proj_bias', ctypes.c_void_p),
                ('out_proj_weight', ctypes.c_void_p), ('out_proj_bias', ctypes.c_void_p),
                ('enc_pe', ctypes.c_void_p), ('num_layers', ctypes.c_int),
                ('layers', _ROLayer * RO_MAX_LAYERS)]


SFB_TR_NONE, SFB_TR_TRANSFORMER, SFB_TR_MLP = 0, 1, 2
TR_MAX_LAYERS = 4
TR_MLP_KEYS = ('ln.weight', 'ln.bias', 'mlp.0.weight', 'mlp.0.bias', 'mlp.2.weight', 'mlp.2.bias')
TR_RNN_KEYS = ('rnn.weight_ih_l0', 'rnn.weight_hh_l0', 'rnn.bias_ih_l0', 'rnn.bias_hh_l0',
               'out_projector.weight', 'out_projector.bias')
TR_KD_KEYS = ('kernel_dist.0.weight', 'kernel_dist.0.bias', 'kernel_dist.1.weight', 'kernel_dist.1.bias',
              'kernel_dist.3.weight', 'kernel_dist.3.bias')


class _TRWeights(ctypes.Structure):
    """sfb_tr_weights (include/sfb200.h): same field order."""
    _fields_ = ([('pred_type', ctypes.c_int), ('num_layers', ctypes.c_int), ('num_heads', ctypes.c_int),
                 ('ffn_dim', ctypes.c_int), ('norm_first', ctypes.c_int), ('layers', _ROLayer * TR_MAX_LAYERS),
                 ('mlp_hidden', ctypes.c_int)]
                + [(k.replace('.', '_'), ctypes.c_void_p) for k in TR_MLP_KEYS]
                + [('rnn_hidden', ctypes.c_int)]
                + [(k.replace('.', '_'), ctypes.c_void_p) for k in TR_RNN_KEYS]
                + [('kernel_mlp', ctypes.c_int)]
                + [(k.replace('.', '_'), ctypes.c_void_p) for k in TR_KD_KEYS])


def lib_path():
    return _LIB_PATH


def _bind(path, debug):
    lib = ctypes.CDLL(path)
    c = ctypes
    lib.sfb_version.restype = c.c_int
    lib.sfb_strerror.restype = c.c_char_p
    lib.sfb_strerror.argtypes = [c.c_int]
    lib.sfb_launch_count.restype = c.c_longlong
    if debug:
        lib.sfb_debug_set_profile.restype = None
        lib.sfb_debug_set_profile.argtypes = [c.c_void_p, c.c_int]
        lib.sfb_debug_umma_gemm.restype = c.c_int
        lib.sfb_debug_umma_gemm.argtypes = [c.c_void_p, c.c_void_p, c.c_void_p, c.c_int, c.c_int, c.c_int,
                                            c.c_void_p, c.c_size_t, c.c_void_p]
    lib.sfb_sa_workspace_bytes.restype = c.c_size_t
    lib.sfb_sa_workspace_bytes.argtypes = [c.c_int] * 7
    lib.sfb_sa_prepare.restype = c.c_int
    lib.sfb_sa_prepare.argtypes = [c.POINTER(_SAWeights), c.c_int, c.c_int, c.c_int, c.c_void_p, c.c_size_t,
                                   c.c_void_p]
    lib.sfb_sa_forward.restype = c.c_int
    lib.sfb_sa_forward.argtypes = [
        c.c_void_p, c.c_int, c.c_int64, c.c_void_p, c.c_void_p, c.c_void_p,
        c.POINTER(_SAWeights), c.c_int, c.c_int, c.c_int, c.c_int, c.c_int, c.c_int, c.c_int,
        c.c_float, c.c_int, c.c_int, c.c_uint, c.c_void_p, c.c_size_t, c.c_void_p]
    lib.sfb_enc_tail_workspace_bytes.restype = c.c_size_t
    lib.sfb_enc_tail_workspace_bytes.argtypes = [c.c_int]
    lib.sfb_enc_tail_tiles_bytes.restype = c.c_size_t
    lib.sfb_enc_tail_tiles_bytes.argtypes = [c.c_int, c.c_int, c.c_int]
    lib.sfb_enc_tail_prepare.restype = c.c_int
    lib.sfb_enc_tail_prepare.argtypes = [c.POINTER(_EncTailWeights), c.c_int, c.c_void_p, c.c_size_t, c.c_void_p]
    lib.sfb_enc_tail_forward.restype = c.c_int
    lib.sfb_enc_tail_forward.argtypes = [c.c_void_p, c.c_int64, c.c_int, c.c_int, c.c_int, c.c_int, c.c_void_p,
                                         c.c_size_t, c.c_void_p, c.c_size_t, c.c_int, c.c_uint, c.c_void_p]
    lib.sfb_rollout_workspace_bytes.restype = c.c_size_t
    lib.sfb_rollout_workspace_bytes.argtypes = [c.c_int, c.c_int, c.c_int, c.c_int]
    lib.sfb_rollout_prepare.restype = c.c_int
    lib.sfb_rollout_prepare.argtypes = [c.POINTER(_ROWeights), c.c_int, c.c_int, c.c_int,
                                        c.c_void_p, c.c_size_t, c.c_void_p]
    lib.sfb_rollout_forward.restype = c.c_int
    lib.sfb_rollout_forward.argtypes = [
        c.c_void_p, c.c_void_p, c.POINTER(_ROWeights), c.c_int, c.c_int, c.c_int, c.c_int,
        c.c_int, c.c_int, c.c_int, c.c_int, c.c_int, c.c_int, c.c_uint, c.c_void_p, c.c_size_t,
        c.c_void_p]
    lib.sfb_decode_combine.restype = c.c_int
    lib.sfb_decode_combine.argtypes = [c.c_void_p, c.c_void_p, c.c_void_p, c.c_void_p, c.c_void_p, c.c_int, c.c_int,
                                       c.c_int, c.c_float, c.c_void_p]
    lib.sfb_postproc_mask.restype = c.c_int
    lib.sfb_postproc_mask.argtypes = [c.c_void_p, c.c_void_p, c.c_void_p, c.c_int, c.c_int, c.c_int, c.c_float,
                                      c.c_void_p]
    lib.sfb_transition_workspace_bytes.restype = c.c_size_t
    lib.sfb_transition_workspace_bytes.argtypes = [c.POINTER(_TRWeights), c.c_int]
    lib.sfb_transition_prepare.restype = c.c_int
    lib.sfb_transition_prepare.argtypes = [c.POINTER(_TRWeights), c.c_int, c.c_void_p, c.c_size_t, c.c_void_p]
    lib.sfb_transition_forward.restype = c.c_int
    lib.sfb_transition_forward.argtypes = [c.POINTER(_TRWeights), c.c_int, c.c_int, c.c_int, c.c_void_p, c.c_longlong,
                                           c.c_int, c.c_void_p, c.c_void_p, c.c_void_p, c.c_void_p, c.c_void_p,
                                           c.c_void_p, c.c_void_p, c.c_void_p, c.c_size_t, c.c_void_p]
    return lib


def load():
    """Load libsfb200.so (once).  Raises SfbError if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            raise SfbError(f'{_LIB_PATH} is missing: run `python -m slotformer_b200.build` '
                           '(there is no fallback path)')
        _lib = _bind(_LIB_PATH, debug=False)
    return _lib


def load_debug():
    """Load libsfb200_debug.so (-DSFB_DEBUG: tcgen05 self-test, timeline hook, SFB_DBG kernel switches); used by
    tests/test_gpu_umma.py and scripts/prof_*.py only."""
    global _lib_debug
    if _lib_debug is None:
        if not os.path.exists(_LIB_DEBUG_PATH):
            raise SfbError(f'{_LIB_DEBUG_PATH} is missing: run `python -m slotformer_b200.build --debug`')
        _lib_debug = _bind(_LIB_DEBUG_PATH, debug=True)
    return _lib_debug


def use_debug_library():
    """Route every engine call of this process through the debug build (profiling scripts)."""
    global _lib
    _lib = load_debug()
    return _lib


def exported_symbols(debug=False):
    """Names declared in include/sfb200.h (used by the CPU-side ABI test)."""
    names = ['sfb_version', 'sfb_strerror', 'sfb_launch_count', 'sfb_sa_workspace_bytes', 'sfb_sa_prepare',
             'sfb_sa_forward', 'sfb_enc_tail_workspace_bytes', 'sfb_enc_tail_tiles_bytes', 'sfb_enc_tail_prepare',
             'sfb_enc_tail_forward', 'sfb_rollout_workspace_bytes', 'sfb_rollout_prepare',
             'sfb_rollout_forward', 'sfb_decode_combine', 'sfb_postproc_mask', 'sfb_transition_workspace_bytes',
             'sfb_transition_prepare', 'sfb_transition_forward']
    if debug:
        names += ['sfb_debug_set_profile', 'sfb_debug_umma_gemm']
    return names


def umma_gemm(W, X):
    """Self-test hook (debug build): X [N,K] @ W[M,K]^T on the tcgen05 path (fp16 operands, fp32 accumulate)."""
    lib = load_debug()
    _require_cuda_f32('W', W)
    _require_cuda_f32('X', X)
    M, K = W.shape
    N = X.shape[0]
    out = torch.empty((N, M), dtype=torch.float32, device=W.device)
    ws = torch.empty(M * K * 2, dtype=torch.uint8, device=W.device)
    with torch.cuda.device(W.device):
        _check(lib.sfb_debug_umma_gemm(W.contiguous().data_ptr(), X.contiguous().data_ptr(), out.data_ptr(),
                                       M, N, K, ws.data_ptr(), ws.numel(), _stream(W.device)))
    return out


def launch_count():
    return int(load().sfb_launch_count())


class HotPathPipeline:
    """Two-stage software pipeline over consecutive clip batches.

    Slot Attention of batch i+1 (HBM-bound, persistent CTAs) and the rollout of batch i (latency-bound, one
    CTA per clip) use different resources, so they run concurrently on two streams: the rollout keeps
    ``clips`` SMs, Slot Attention is capped at the remaining ones.  ``submit`` returns at once; the returned
    event completes when that batch's predicted slots are ready.  Results are identical to calling the two
    modules back to back (same kernels, same order per batch).
    """

    def __init__(self, slot_attention, rollouter, device, clips):
        self.sa, self.ro, self.device = slot_attention, rollouter, torch.device(device)
        self.s_sa = torch.cuda.Stream(self.device)
        self.s_ro = torch.cuda.Stream(self.device)
        sms = torch.cuda.get_device_properties(self.device).multi_processor_count
        self.sa_ctas = sms - clips if 0 < clips <= sms // 2 else 0     # 0: no partition, plain stream order

    def __enter__(self):
        # the CTA cap is an argument of every sfb_sa_forward call made through this module while the pipeline is open
        self._saved_ctas = self.sa.max_ctas
        self.sa.max_ctas = self.sa_ctas
        cur = torch.cuda.current_stream(self.device)
        self.s_sa.wait_stream(cur)
        self.s_ro.wait_stream(cur)
        return self

    def __exit__(self, *exc):
        cur = torch.cuda.current_stream(self.device)
        cur.wait_stream(self.s_sa)
        cur.wait_stream(self.s_ro)
        self.sa.max_ctas = self._saved_ctas
        return False

    def capture(self, batches, clips, frames_per_clip, pred_len):
        """Capture ``submit`` of every (feats, init_slots) pair in ``batches`` into ONE CUDA graph (both streams, with
        their cross-stream dependencies), so that replaying it costs the host a single launch: the overlap of the two
        stages no longer depends on the host staying ahead of the GPU.  Returns (graph, [(slots, pred), ...]); the
        outputs are static tensors that every ``graph.replay()`` overwrites; inputs are read in place."""
        dev = self.device
        cap = torch.cuda.Stream(dev)
        cap.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(cap):                         # warm-up outside the capture
            with self:
                for f, s0 in batches[:2]:
                    self.submit(f, s0, clips, frames_per_clip, pred_len)
        torch.cuda.current_stream(dev).wait_stream(cap)
        torch.cuda.synchronize(dev)
        graph = torch.cuda.CUDAGraph()
        outs = []
        with torch.cuda.graph(graph, stream=cap):
            with self:
                for f, s0 in batches:
                    sl, pr, _ = self.submit(f, s0, clips, frames_per_clip, pred_len)
                    outs.append((sl, pr))
        return graph, outs

    def submit(self, feats, init_slots, clips, frames_per_clip, pred_len, after=None, timing=None):
        """feats [clips*frames_per_clip, N, C], init_slots [.., K, D] -> (slots, pred, done_event).

        ``after``: optional event the Slot Attention stage waits for (e.g. the H2D copy of this batch).
        ``timing``: optional list; receives (sa_start, sa_end, ro_start, ro_end) timing events."""
        tm = timing is not None
        with torch.cuda.stream(self.s_sa):
            if after is not None:
                self.s_sa.wait_event(after)
            if tm:
                e0 = torch.cuda.Event(enable_timing=True)
                e0.record(self.s_sa)
            slots = self.sa(feats, init_slots)
            ready = torch.cuda.Event(enable_timing=tm)
            ready.record(self.s_sa)
        if not torch.cuda.is_current_stream_capturing():
            slots.record_stream(self.s_ro)      # (a captured graph owns its memory pool; nothing to record there)
        self.s_ro.wait_event(ready)
        with torch.cuda.stream(self.s_ro):
            if tm:
                e2 = torch.cuda.Event(enable_timing=True)
                e2.record(self.s_ro)
            K, D = slots.shape[1], slots.shape[2]
            pred = self.ro(slots.view(clips, frames_per_clip, K, D), pred_len)
            done = torch.cuda.Event(enable_timing=tm)
            done.record(self.s_ro)
        if tm:
            timing.append((e0, ready, e2, done))
        return slots, pred, done


def _check(code):
    if code != 0:
        raise SfbError(f'libsfb200: {load().sfb_strerror(code).decode()} (code {code})')


def _stream(device):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _require_cuda_f32(name, t):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise SfbError(f'{name} must be a CUDA tensor (the engine has no CPU path)')
    if t.dtype != torch.float32:
        raise SfbError(f'{name} must be float32, got {t.dtype}')


def _weights_key(tensors):
    return tuple((t.data_ptr(), t._version) for t in tensors)


# --------------------------------------------------------------------------- #
# Encoder tail
# --------------------------------------------------------------------------- #
class FeatureTiles:
    """LayerNorm-ed per-pixel features as the fp16 operand tiles of the tcgen05 Slot Attention passes
    (sfb_enc_tail_forward output, SFB_DTYPE_TILES16): stands in for the reference's fp32 [..., N, C] feature grid
    wherever it is only indexed over its leading (frame) dimensions and handed to ``SlotAttention``.
    ``data``: fp16 [*lead, halves_per_frame]; one frame = ceil-to-chunk(N) / 128 tiles of 32 KB."""

    def __init__(self, data, N, C):
        self.data, self.N, self.C = data, int(N), int(C)

    shape = property(lambda self: tuple(self.data.shape[:-1]) + (self.N, self.C))
    device = property(lambda self: self.data.device)
    is_cuda = property(lambda self: self.data.is_cuda)
    dtype = torch.float16
    requires_grad = False

    def dim(self):
        return self.data.dim() + 1

    def __getitem__(self, idx):
        return FeatureTiles(self.data[idx], self.N, self.C)

    def unflatten(self, dim, sizes):
        assert dim == 0
        return FeatureTiles(self.data.unflatten(0, sizes), self.N, self.C)

    def detach(self):
        return self

    def contiguous(self):
        return FeatureTiles(self.data.contiguous(), self.N, self.C)

    def to_dense(self):
        """[*lead, N, C] fp32 values of t (for tests): undoes the tile / panel / 128-byte-swizzle layout."""
        lead = self.data.shape[:-1]
        d = self.data.reshape(-1, self.data.shape[-1] // 16384, 2, 128, 8, 8)     # frame, tile, panel, row, chunk, 8 halves
        rows = torch.arange(128, device=d.device)
        chunk = torch.arange(8, device=d.device)[None, :] ^ (rows[:, None] & 7)         # logical chunk -> stored chunk
        d = torch.gather(d, 4, chunk[None, None, None, :, :, None].expand(d.shape[0], d.shape[1], 2, 128, 8, 8))
        d = d.permute(0, 1, 3, 2, 4, 5).reshape(d.shape[0], -1, 128)                # frame, pixel, channel
        return d[:, :self.N].float().reshape(*lead, self.N, self.C)


class EncoderTailEngine:
    """Launcher state for sfb_enc_tail_prepare / sfb_enc_tail_forward."""

    def __init__(self):
        self._ws = None
        self._key = None
        self._captured = []

    def invalidate(self):
        self._key = None

    def forward(self, cnn_out, weights, C, max_ctas=0):
        """cnn_out [F, 64, H, W] f32 (CNN encoder output; NCHW-contiguous or torch.channels_last memory, read as it
        is) -> FeatureTiles [F] (N = H*W, C features)."""
        lib = load()
        _require_cuda_f32('cnn_out', cnn_out)
        if cnn_out.dim() != 4 or cnn_out.shape[1] != 64:
            raise SfbError(f'cnn_out must be [frames, 64, H, W], got {tuple(cnn_out.shape)}')
        F_, _, H, W = cnn_out.shape
        # channels-last memory (what cuDNN's tensor-core convolutions write) is consumed directly: no NCHW copy
        nhwc = (H * W > 1 and not cnn_out.is_contiguous()
                and cnn_out[:1].is_contiguous(memory_format=torch.channels_last) and cnn_out.stride(0) >= 64 * H * W)
        if not nhwc:
            cnn_out = cnn_out.contiguous()
        dev = cnn_out.device
        per_frame = int(lib.sfb_enc_tail_tiles_bytes(1, H * W, C))
        ws_bytes = int(lib.sfb_enc_tail_workspace_bytes(C))
        if per_frame == 0 or ws_bytes == 0:
            raise SfbError(f'unsupported encoder-tail shape C={C} H={H} W={W}')
        tiles = torch.empty((F_, per_frame // 2), dtype=torch.float16, device=dev)
        if F_ == 0:
            return FeatureTiles(tiles, H * W, C)
        cw = _EncTailWeights()
        wt = []
        for k in ENC_TAIL_KEYS:
            t = weights[k]
            _require_cuda_f32(k, t)
            t = t if t.is_contiguous() else t.contiguous()
            wt.append(t)
            setattr(cw, k.replace('.', '_'), t.data_ptr())
        with torch.cuda.device(dev):
            if self._ws is None or self._ws.device != dev:
                self._ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
                self._key = None
            if torch.cuda.is_current_stream_capturing() and not any(w is self._ws for w in self._captured):
                self._captured.append(self._ws)
            key = _weights_key(wt)
            if key != self._key:
                _check(lib.sfb_enc_tail_prepare(ctypes.byref(cw), C, self._ws.data_ptr(), ws_bytes, _stream(dev)))
                self._key = key
            _check(lib.sfb_enc_tail_forward(cnn_out.data_ptr(), cnn_out.stride(0), F_, H, W, C, tiles.data_ptr(),
                                            tiles.numel() * 2, self._ws.data_ptr(), ws_bytes, int(max_ctas),
                                            SFB_ET_NHWC if nhwc else 0, _stream(dev)))
        return FeatureTiles(tiles, H * W, C)


# --------------------------------------------------------------------------- #
# Slot Attention
# --------------------------------------------------------------------------- #
class SlotAttentionEngine:
    """Per-module launcher state (workspace) for sfb_sa_forward."""

    def __init__(self):
        self._ws = None
        self._key = None
        self._captured = []      # workspaces referenced by live CUDA graphs: never freed under them

    def invalidate(self):
        """Forget the folded weights (call after in-place weight updates that bypass autograd's version counter,
        e.g. ``p.data.copy_``): the next call runs sfb_sa_prepare again."""
        self._key = None

    def forward(self, feats, slots, weights, num_iterations, eps, mlp_hidden_size,
                return_mask=False, chunk_frames=0, max_ctas=0, flags=0):
        """feats [B,N,C] f32 (rows contiguous; batch stride free), slots [B,K,D] f32.

        ``weights``: dict state_dict-key -> CUDA f32 tensor (SA_WEIGHT_KEYS).
        Returns slots [B,K,D] (and seg mask [B,K,N] if ``return_mask``).
        """
        lib = load()
        tiles = isinstance(feats, FeatureTiles)
        if tiles:
            if not feats.is_cuda:
                raise SfbError('feature tiles must live on a CUDA device')
            feat_dtype = SFB_DTYPE_TILES16
        elif isinstance(feats, torch.Tensor) and feats.dtype == torch.bfloat16 and feats.is_cuda:
            feat_dtype = SFB_DTYPE_BF16
        else:
            _require_cuda_f32('inputs', feats)
            feat_dtype = SFB_DTYPE_F32
        _require_cuda_f32('slots', slots)
        if feats.dim() != 3 or slots.dim() != 3 or feats.shape[0] != slots.shape[0]:
            raise SfbError(f'bad shapes: inputs {tuple(feats.shape)}, slots {tuple(slots.shape)}')
        B, N, C = feats.shape
        K, D = slots.shape[1], slots.shape[2]
        if tiles:
            raw = feats.data if feats.data.stride(-1) == 1 else feats.data.contiguous()
            bstride = raw.stride(0) if B > 1 else raw.shape[-1]
            feats = raw
        else:
            if feats.stride(2) != 1 or feats.stride(1) != C:
                feats = feats.contiguous()
            bstride = feats.stride(0) if B > 1 else N * C
        slots = slots.contiguous()
        dev = feats.device
        if B == 0:
            out = slots.new_zeros((0, K, D))
            return (out, slots.new_zeros((0, K, N))) if return_mask else out
        ws_bytes = int(lib.sfb_sa_workspace_bytes(B, N, C, D, int(mlp_hidden_size),
                                                  1 if tiles else int(num_iterations), int(chunk_frames)))
        if ws_bytes == 0:
            raise SfbError(f'unsupported Slot Attention shape B={B} N={N} C={C} D={D}')
        if self._ws is None or self._ws.device != dev or self._ws.numel() < ws_bytes:
            self._ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
            self._key = None
        if torch.cuda.is_current_stream_capturing() and not any(w is self._ws for w in self._captured):
            self._captured.append(self._ws)      # a graph now holds this pointer: keep it alive after a realloc
        wt = []
        cw = _SAWeights()
        for k in SA_WEIGHT_KEYS:
            t = weights[k]
            _require_cuda_f32(k, t)
            t = t if t.is_contiguous() else t.contiguous()
            wt.append(t)
            setattr(cw, k.replace('.', '_'), t.data_ptr())
        out = torch.empty((B, K, D), dtype=torch.float32, device=dev)
        mask = torch.empty((B, K, N), dtype=torch.float32, device=dev) if return_mask else None
        with torch.cuda.device(dev):
            key = _weights_key(wt)
            if key != self._key:       # weights changed (or first call): fold + pack once
                _check(lib.sfb_sa_prepare(ctypes.byref(cw), C, D, int(mlp_hidden_size), self._ws.data_ptr(),
                                          self._ws.numel(), _stream(dev)))
                self._key = key
            rc = lib.sfb_sa_forward(
                feats.data_ptr(), feat_dtype, bstride, slots.data_ptr(), out.data_ptr(),
                mask.data_ptr() if return_mask else None, ctypes.byref(cw), B, N, C, D,
                int(mlp_hidden_size), K, int(num_iterations), float(eps), int(chunk_frames),
                int(max_ctas), int(flags), self._ws.data_ptr(), ws_bytes, _stream(dev))
        _check(rc)
        if return_mask:
            return out, mask
        return out


# --------------------------------------------------------------------------- #
# Rollout
# --------------------------------------------------------------------------- #
class RolloutEngine:
    """Per-module launcher state for sfb_rollout_prepare / sfb_rollout_forward."""

    def __init__(self):
        self._ws = None
        self._key = None
        self._captured = []

    def invalidate(self):
        self._key = None

    def forward(self, hist, weights, enc_pe, num_layers, num_heads, pred_len, mode='slide',
                cond_len=0, flags=0):
        """hist [B,T_h,K,Ds] f32 -> [B,pred_len,K,Ds] f32.

        ``weights``: dict of SlotRollouter state_dict keys -> CUDA f32 tensors.
        ``enc_pe``: per-token positional table [pe_frames*K, d] f32.
        """
        lib = load()
        _require_cuda_f32('x', hist)
        if hist.dim() != 4:
            raise SfbError(f'x must be [B, T, K, Ds], got {tuple(hist.shape)}')
        hist = hist.contiguous()
        B, T_h, K, Ds = hist.shape
        dev = hist.device
        if B == 0 or pred_len == 0:
            return hist.new_zeros((B, pred_len, K, Ds))
        d = weights['in_proj.weight'].shape[0]
        F = weights['transformer_encoder.layers.0.linear1.weight'].shape[0]
        cw = _ROWeights()
        keep = []

        def ptr(name):
            t = weights[name]
            _require_cuda_f32(name, t)
            t = t if t.is_contiguous() else t.contiguous()
            keep.append(t)
            return t.data_ptr()

        cw.in_proj_weight = ptr('in_proj.weight')
        cw.in_proj_bias = ptr('in_proj.bias')
        cw.out_proj_weight = ptr('out_proj.weight')
        cw.out_proj_bias = ptr('out_proj.bias')
        _require_cuda_f32('enc_pe', enc_pe)
        enc_pe = enc_pe.contiguous()
        cw.enc_pe = enc_pe.data_ptr()
        cw.num_layers = int(num_layers)
        if num_layers > RO_MAX_LAYERS:
            raise SfbError(f'num_layers {num_layers} > {RO_MAX_LAYERS}')
        for i in range(num_layers):
            for k in RO_LAYER_KEYS:
                setattr(cw.layers[i], k.replace('.', '_'),
                        ptr(f'transformer_encoder.layers.{i}.{k}'))
        ws_bytes = int(lib.sfb_rollout_workspace_bytes(Ds, d, F, num_layers))
        out = torch.empty((B, pred_len, K, Ds), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            if self._ws is None or self._ws.device != dev or self._ws.numel() < ws_bytes:
                self._ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
                self._key = None
            if torch.cuda.is_current_stream_capturing() and not any(w is self._ws for w in self._captured):
                self._captured.append(self._ws)
            key = _weights_key(keep)
            if key != self._key:
                _check(lib.sfb_rollout_prepare(ctypes.byref(cw), Ds, d, F, self._ws.data_ptr(),
                                               ws_bytes, _stream(dev)))
                self._key = key
            rc = lib.sfb_rollout_forward(
                hist.data_ptr(), out.data_ptr(), ctypes.byref(cw), B, T_h, K, Ds, d, F,
                int(num_heads), int(pred_len), SFB_RO_GROW if mode == 'grow' else SFB_RO_SLIDE,
                int(cond_len or 0), int(flags), self._ws.data_ptr(), ws_bytes, _stream(dev))
        _check(rc)
        return out


# --------------------------------------------------------------------------- #
# SAVi slot transition
# --------------------------------------------------------------------------- #
class TransitionEngine:
    """Per-module launcher state for sfb_transition_prepare / sfb_transition_forward: predictor ->
    kernel_dist_layer -> sample of reference StoSAVi.encode (savi.py:393-410) as one kernel launch per frame."""

    def __init__(self):
        self._ws = None
        self._key = None
        self._captured = []

    def invalidate(self):
        self._key = None

    @staticmethod
    def _pack(spec):
        """spec -> (_TRWeights, tensors kept alive).  ``spec``: dict(pred_type, num_layers, num_heads, ffn_dim,
        norm_first, mlp_hidden, rnn_hidden, kernel_mlp, weights={name: CUDA f32 tensor})."""
        cw = _TRWeights()
        keep = []
        W = spec['weights']

        def ptr(name):
            t = W[name]
            _require_cuda_f32(name, t)
            t = t if t.is_contiguous() else t.contiguous()
            keep.append(t)
            return t.data_ptr()

        cw.pred_type = int(spec['pred_type'])
        cw.num_layers = int(spec.get('num_layers', 0))
        cw.num_heads = int(spec.get('num_heads', 0))
        cw.ffn_dim = int(spec.get('ffn_dim', 0))
        cw.norm_first = int(bool(spec.get('norm_first', True)))
        cw.mlp_hidden = int(spec.get('mlp_hidden', 0))
        cw.rnn_hidden = int(spec.get('rnn_hidden', 0))
        cw.kernel_mlp = int(bool(spec.get('kernel_mlp', True)))
        if cw.pred_type == SFB_TR_TRANSFORMER:
            if not 1 <= cw.num_layers <= TR_MAX_LAYERS:
                raise SfbError(f'transition: {cw.num_layers} predictor layers (max {TR_MAX_LAYERS})')
            for i in range(cw.num_layers):
                for k in RO_LAYER_KEYS:
                    setattr(cw.layers[i], k.replace('.', '_'), ptr(f'layers.{i}.{k}'))
        elif cw.pred_type == SFB_TR_MLP:
            for k in TR_MLP_KEYS:
                setattr(cw, k.replace('.', '_'), ptr(k))
        if cw.rnn_hidden > 0:
            for k in TR_RNN_KEYS:
                setattr(cw, k.replace('.', '_'), ptr(k))
        for k in (TR_KD_KEYS if cw.kernel_mlp else TR_KD_KEYS[:2]):
            setattr(cw, k.replace('.', '_'), ptr(k))
        return cw, keep

    @staticmethod
    def supported(spec, D):
        """True if the kernel covers this structure (layer widths <= 1024, multiples of 32, ...)."""
        try:
            cw, _ = TransitionEngine._pack(spec)
        except (SfbError, KeyError):
            return False
        return int(load().sfb_transition_workspace_bytes(ctypes.byref(cw), int(D))) > 0

    def forward(self, spec, prev, use_predictor, B, state=None, noise=None):
        """prev: [B, K, D] previous slots, or [1, K, D] rows shared by every clip (init latents, use_predictor False).
        state: None or (h, c) with B*K rows of rnn_hidden.  Returns (dist [B,K,2D], slots0 [B,K,D], new_state)."""
        lib = load()
        _require_cuda_f32('prev', prev)
        if prev.dim() != 3 or prev.shape[0] not in (1, B):
            raise SfbError(f'prev must be [B, K, D] or [1, K, D], got {tuple(prev.shape)}')
        prev = prev.contiguous()
        K, D = prev.shape[1], prev.shape[2]
        dev = prev.device
        cstride = K * D if (prev.shape[0] == B and B > 1) else 0
        cw, keep = self._pack(spec)
        ws_bytes = int(lib.sfb_transition_workspace_bytes(ctypes.byref(cw), D))
        if ws_bytes == 0 or K > 8:
            raise SfbError(f'unsupported transition structure (D={D}, K={K})')
        H = cw.rnn_hidden
        rnn = bool(use_predictor) and H > 0
        h_in = c_in = h_out = c_out = None
        if rnn:
            if state is not None:
                h_in, c_in = (t.reshape(B * K, H).contiguous() for t in state)
                _require_cuda_f32('h', h_in)
                _require_cuda_f32('c', c_in)
            h_out = torch.empty((B * K, H), dtype=torch.float32, device=dev)
            c_out = torch.empty((B * K, H), dtype=torch.float32, device=dev)
        if noise is not None:
            _require_cuda_f32('noise', noise)
            noise = noise.contiguous()
        dist = torch.empty((B, K, 2 * D), dtype=torch.float32, device=dev)
        slots = torch.empty((B, K, D), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            if self._ws is None or self._ws.device != dev or self._ws.numel() < ws_bytes:
                self._ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
                self._key = None
            if torch.cuda.is_current_stream_capturing() and not any(w is self._ws for w in self._captured):
                self._captured.append(self._ws)
            key = _weights_key(keep)
            if key != self._key:
                _check(lib.sfb_transition_prepare(ctypes.byref(cw), D, self._ws.data_ptr(), ws_bytes, _stream(dev)))
                self._key = key
            p = lambda t: None if t is None else t.data_ptr()
            _check(lib.sfb_transition_forward(ctypes.byref(cw), D, B, K, prev.data_ptr(), cstride, int(bool(use_predictor)),
                                              p(h_in), p(c_in), p(noise), dist.data_ptr(), slots.data_ptr(), p(h_out),
                                              p(c_out), self._ws.data_ptr(), ws_bytes, _stream(dev)))
        return dist, slots, ((h_out, c_out) if rnn else None)


# --------------------------------------------------------------------------- #
# Decoder epilogue
# --------------------------------------------------------------------------- #
FG_THRE = 0.5      # reference video_prediction/vp_utils.py:11


def decode_combine(dec_out, want_seg=False, fg_thre=FG_THRE):
    """Tail of StoSAVi.decode (reference savi.py:519-523) as one streaming kernel.

    dec_out [B, K, 4, H, W] f32 (deconv output) -> (recon_combined [B,3,H,W], masks [B,K,1,H,W]) and, with
    ``want_seg``, the post-processed segmentation [B,H,W] int64 (reference postproc_mask)."""
    lib = load()
    _require_cuda_f32('dec_out', dec_out)
    if dec_out.dim() != 5 or dec_out.shape[2] != 4:
        raise SfbError(f'dec_out must be [B, K, 4, H, W], got {tuple(dec_out.shape)}')
    dec_out = dec_out.contiguous()
    B, K, _, H, W = dec_out.shape
    dev = dec_out.device
    masks = torch.empty((B, K, 1, H, W), dtype=torch.float32, device=dev)
    recon = torch.empty((B, 3, H, W), dtype=torch.float32, device=dev)
    seg = torch.empty((B, H, W), dtype=torch.int64, device=dev) if want_seg else None
    ws = torch.empty((max(B * K, 1),), dtype=torch.int32, device=dev) if want_seg else None
    if B > 0:
        with torch.cuda.device(dev):
            _check(lib.sfb_decode_combine(dec_out.data_ptr(), masks.data_ptr(), recon.data_ptr(),
                                          seg.data_ptr() if want_seg else None, ws.data_ptr() if want_seg else None,
                                          B, K, H * W, float(fg_thre), _stream(dev)))
    return (recon, masks, seg) if want_seg else (recon, masks)


def postproc_mask(batch_masks, fg_thre=FG_THRE):
    """Reference postproc_mask (vp_utils.py:20-41): batch_masks [B, T, N, 1, H, W] f32 -> [B, T, H, W] int64."""
    lib = load()
    _require_cuda_f32('batch_masks', batch_masks)
    if batch_masks.dim() != 6 or batch_masks.shape[3] != 1:
        raise SfbError(f'batch_masks must be [B, T, N, 1, H, W], got {tuple(batch_masks.shape)}')
    B, T, N, _, H, W = batch_masks.shape
    m = batch_masks.contiguous()
    dev = m.device
    seg = torch.empty((B, T, H, W), dtype=torch.int64, device=dev)
    if B * T > 0:
        ws = torch.empty((B * T * N,), dtype=torch.int32, device=dev)
        with torch.cuda.device(dev):
            _check(lib.sfb_postproc_mask(m.data_ptr(), seg.data_ptr(), ws.data_ptr(), B * T, N, H * W,
                                         float(fg_thre), _stream(dev)))
    return seg
