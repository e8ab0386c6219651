"""CPU oracle for the two SlotFormer hot paths -- TEST INFRASTRUCTURE ONLY.

This file is a plain-numpy restatement of the reference algorithm.  It is the
checker for the CUDA path; it is never imported by ``slotformer_b200`` (the
product) -- only by ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py``.

Parity status: PINNED.  The reference ships no golden vectors for this path
(SURVEY.md section 4), so the oracle is pinned against outputs of the
*reference itself*, imported unmodified from /root/reference in the build
container by ``tests/golden/make_golden.py`` (committed, with the vectors it
produced in ``tests/golden/*.npz``).  ``tests/test_oracle_golden.py`` checks
every function here against those vectors.

Reference lines followed (paths relative to /root/reference/slotformer):
  * slot_attention      base_slots/models/savi.py:56-102
  * seg-mask variant    base_slots/models/steve.py:43-73 (mask taken at :54-55)
  * sin_pos_enc         video_prediction/models/slotformer.py:10-16
  * rollout (slide)     video_prediction/models/slotformer.py:85-126
  * rollout (grow)      video_prediction/models/single_step_slotformer.py:49-90
  * encoder layer       torch.nn.TransformerEncoderLayer(norm_first=True,
                        batch_first=True, activation=relu), eval mode, as built
                        at slotformer.py:72-80
  * GRU cell            torch.nn.GRUCell gate order (r, z, n)
  * transition          base_slots/models/savi.py:394-403 with predictor.py:20-113 (Transformer / residual-MLP predictor,
                        LSTM wrapper, torch.nn.LSTM gate order i, f, g, o), kernel_dist_layer savi.py:200-212,
                        _sample_dist savi.py:355-363 (pinned by tests/golden/transition.npz)
  * decode_combine      base_slots/models/savi.py:519-523 (pinned by tests/golden/decode.npz)
  * postproc_mask       video_prediction/vp_utils.py:20-41 (pinned by tests/golden/decode.npz)

Weights are dicts keyed by the reference ``state_dict`` names, so the same
dict loads into the reference modules (that is how the goldens are made).
"""
import numpy as np

LN_EPS = 1e-5  # torch.nn.LayerNorm default, used everywhere in the reference


# --------------------------------------------------------------------------- #
# helpers
# --------------------------------------------------------------------------- #
def to_bf16(x):
    """Round-to-nearest-even to bfloat16, returned in the input's float type.

    Used only by the ``operand_round`` hook below to emulate the tensor-core
    operand precision of the CUDA rollout kernel.
    """
    a = np.ascontiguousarray(x, dtype=np.float32)
    u = a.view(np.uint32).astype(np.uint64)
    rounded = ((u + 0x7FFF + ((u >> 16) & 1)) >> 16) << 16
    out = rounded.astype(np.uint32).view(np.float32)
    return out.astype(x.dtype if hasattr(x, 'dtype') else np.float32)


def to_tf32(x):
    """Truncate fp32 mantissa to 10 bits (tensor-core TF32 operand read)."""
    a = np.ascontiguousarray(x, dtype=np.float32)
    u = a.view(np.uint32) & np.uint32(0xFFFFE000)
    return u.view(np.float32).astype(x.dtype)


def layer_norm(x, weight, bias, eps=LN_EPS):
    mu = x.mean(axis=-1, keepdims=True)
    var = ((x - mu) ** 2).mean(axis=-1, keepdims=True)  # biased, like torch
    return (x - mu) / np.sqrt(var + eps) * weight + bias


def _sigmoid(x):
    return 1.0 / (1.0 + np.exp(-x))


def _softmax(x, axis):
    m = x.max(axis=axis, keepdims=True)
    e = np.exp(x - m)
    return e / e.sum(axis=axis, keepdims=True)


def _cast(w, dtype):
    return {k: np.asarray(v, dtype=dtype) for k, v in w.items()}


# --------------------------------------------------------------------------- #
# hot path 1: Slot Attention   (savi.py:56-102, steve.py:43-73)
# --------------------------------------------------------------------------- #
def gru_cell(x, h, w):
    """torch.nn.GRUCell: gates stacked (r, z, n) in weight_ih / weight_hh."""
    D = h.shape[-1]
    gi = x @ w['gru.weight_ih'].T + w['gru.bias_ih']
    gh = h @ w['gru.weight_hh'].T + w['gru.bias_hh']
    r = _sigmoid(gi[..., :D] + gh[..., :D])
    z = _sigmoid(gi[..., D:2 * D] + gh[..., D:2 * D])
    n = np.tanh(gi[..., 2 * D:] + r * gh[..., 2 * D:])
    return (1.0 - z) * n + z * h


def slot_attention(feats, slots, weights, num_iterations, eps=1e-6,
                   return_mask=False, dtype=np.float64):
    """feats [B,N,C], slots [B,K,D] -> slots [B,K,D] (and seg mask [B,K,N]).

    ``weights`` uses SlotAttention state_dict keys: norm_inputs.{weight,bias},
    project_q.0.{weight,bias}, project_q.1.weight, project_k.weight,
    project_v.weight, gru.{weight_ih,weight_hh,bias_ih,bias_hh},
    mlp.0.{weight,bias}, mlp.1.{weight,bias}, mlp.3.{weight,bias}.
    """
    w = _cast(weights, dtype)
    x = np.asarray(feats, dtype=dtype)
    s = np.asarray(slots, dtype=dtype)
    D = s.shape[-1]
    scale = float(D) ** -0.5                                   # savi.py:35

    xn = layer_norm(x, w['norm_inputs.weight'], w['norm_inputs.bias'])  # :66
    k = xn @ w['project_k.weight'].T                                    # :68
    v = xn @ w['project_v.weight'].T                                    # :70

    seg_mask = None
    for it in range(num_iterations):                                    # :76
        s_prev = s
        q = layer_norm(s, w['project_q.0.weight'], w['project_q.0.bias']) \
            @ w['project_q.1.weight'].T                                 # :80
        logits = scale * np.einsum('bnc,bmc->bnm', k, q)                # :82
        attn = _softmax(logits, axis=-1)                # over slots     :83
        if return_mask and it == num_iterations - 1:    # steve.py:54-55
            seg_mask = attn.transpose(0, 2, 1).copy()
        attn = attn + eps                                               # :87
        attn = attn / attn.sum(axis=1, keepdims=True)   # over pixels    :88
        upd = np.einsum('bnm,bnc->bmc', attn, v)                        # :89
        s = gru_cell(upd, s_prev, w)                                    # :95
        hid = layer_norm(s, w['mlp.0.weight'], w['mlp.0.bias']) \
            @ w['mlp.1.weight'].T + w['mlp.1.bias']
        s = s + np.maximum(hid, 0.0) @ w['mlp.3.weight'].T + w['mlp.3.bias']  # :100
    if return_mask:
        return s, seg_mask
    return s


# --------------------------------------------------------------------------- #
# hot path 2: autoregressive slot-Transformer rollout
# --------------------------------------------------------------------------- #
def sin_pos_enc(seq_len, d_model, dtype=np.float64):
    """[1, seq_len, d_model]; row 0 is the OLDEST frame (position seq_len-1).

    slotformer.py:10-16: halves are concatenated (sin | cos), not interleaved.
    The reference computes this in fp32 torch; we mirror that rounding so the
    table is identical to the ``enc_t_pe`` stored in reference state_dicts.
    """
    inv_freq = (1.0 / (np.float32(10000.0) ** (
        np.arange(0.0, d_model, 2.0, dtype=np.float32) / np.float32(d_model))
    )).astype(np.float32)
    pos = np.arange(seq_len - 1, -1, -1).astype(np.float32)
    ang = np.outer(pos, inv_freq).astype(np.float32)
    pe = np.concatenate([np.sin(ang), np.cos(ang)], axis=-1)
    return pe[None].astype(dtype)


def encoder_layer(h, w, prefix, num_heads, mm):
    """Pre-LN encoder layer, eval mode (dropout off), no mask.

    x = x + out_proj(MHA(LN1(x)));  x = x + W2 relu(W1 LN2(x) + b1) + b2
    ``mm(a, bT)`` computes a @ bT.T and is the hook for operand rounding.
    """
    B, L, d = h.shape
    dh = d // num_heads
    p = prefix
    y = layer_norm(h, w[p + 'norm1.weight'], w[p + 'norm1.bias'])
    qkv = mm(y, w[p + 'self_attn.in_proj_weight']) + w[p + 'self_attn.in_proj_bias']
    q, k, v = np.split(qkv, 3, axis=-1)

    def heads(t):
        return t.reshape(B, L, num_heads, dh).transpose(0, 2, 1, 3)

    q, k, v = heads(q), heads(k), heads(v)
    att = _softmax(np.einsum('bhid,bhjd->bhij', q, k) / np.sqrt(dh), axis=-1)
    o = np.einsum('bhij,bhjd->bhid', att, v).transpose(0, 2, 1, 3).reshape(B, L, d)
    h = h + mm(o, w[p + 'self_attn.out_proj.weight']) + w[p + 'self_attn.out_proj.bias']
    y = layer_norm(h, w[p + 'norm2.weight'], w[p + 'norm2.bias'])
    f = np.maximum(mm(y, w[p + 'linear1.weight']) + w[p + 'linear1.bias'], 0.0)
    return h + mm(f, w[p + 'linear2.weight']) + w[p + 'linear2.bias']


def rollout(hist, weights, pred_len, num_heads, num_layers, mode='slide',
            cond_len=None, dtype=np.float64, operand_round=None,
            return_steps=False):
    """hist [B,T_h,K,Ds] -> [B,pred_len,K,Ds].

    mode='slide': SlotRollouter.forward (slotformer.py:85-126) -- window of
        T_h*K tokens, drop oldest K / append prediction each step.
    mode='grow' : SingleStepSlotRollouter.forward
        (single_step_slotformer.py:49-90) -- T_h must be 1; step s feeds the
        last min(1+s, cond_len)*K tokens with the LAST rows of the PE table.
    ``weights``: SlotRollouter state_dict keys (in_proj.*, transformer_encoder.
        layers.<i>.*, out_proj.*; enc_t_pe optional -- recomputed if absent).
    ``operand_round``: None (exact) or a function applied to both GEMM
        operands (e.g. ``to_bf16``) to emulate tensor-core operand precision.
    """
    w = _cast({k: v for k, v in weights.items()}, dtype)
    x = np.asarray(hist, dtype=dtype)
    B, T_h, K, Ds = x.shape
    d = w['in_proj.weight'].shape[0]

    if operand_round is None:
        def mm(a, bT):
            return a @ bT.T
    else:
        def mm(a, bT):
            return operand_round(a) @ operand_round(bT).T

    if mode == 'slide':
        pe_len = T_h
    else:
        assert mode == 'grow' and T_h == 1 and cond_len is not None
        pe_len = cond_len
    pe_t = w['enc_t_pe'] if 'enc_t_pe' in w else sin_pos_enc(pe_len, d, dtype)
    pe = np.repeat(pe_t[0], K, axis=0)[None]          # [1, pe_len*K, d]  :103-104

    in_x = x.reshape(B, T_h * K, Ds)
    out = []
    for _ in range(pred_len):
        if mode == 'slide':
            win, win_pe = in_x, pe
        else:
            win = in_x[:, -cond_len * K:]              # single_step :79
            win_pe = pe[:, -win.shape[1]:]             # single_step :81
        h = mm(win, w['in_proj.weight']) + w['in_proj.bias'] + win_pe   # :115-117
        for i in range(num_layers):
            h = encoder_layer(h, w, f'transformer_encoder.layers.{i}.',
                              num_heads, mm)                           # :119
        pred = mm(h[:, -K:], w['out_proj.weight']) + w['out_proj.bias']  # :121
        out.append(pred)
        if mode == 'slide':
            in_x = np.concatenate([in_x[:, K:], pred], axis=1)          # :124
        else:
            in_x = np.concatenate([in_x, pred], axis=1)        # single_step :88
    return np.stack(out, axis=1)                                        # :126


# --------------------------------------------------------------------------------------------
# decoder epilogue
# --------------------------------------------------------------------------------------------
def _post_ln_encoder_layer(h, w, prefix, num_heads):
    """torch.nn.TransformerEncoderLayer(norm_first=False): x = LN1(x + SA(x)); x = LN2(x + FF(x))."""
    B, L, d = h.shape
    dh = d // num_heads
    p = prefix
    qkv = h @ w[p + 'self_attn.in_proj_weight'].T + w[p + 'self_attn.in_proj_bias']
    q, k, v = (t.reshape(B, L, num_heads, dh).transpose(0, 2, 1, 3) for t in np.split(qkv, 3, axis=-1))
    att = _softmax(np.einsum('bhid,bhjd->bhij', q, k) / np.sqrt(dh), axis=-1)
    o = np.einsum('bhij,bhjd->bhid', att, v).transpose(0, 2, 1, 3).reshape(B, L, d)
    h = layer_norm(h + o @ w[p + 'self_attn.out_proj.weight'].T + w[p + 'self_attn.out_proj.bias'],
                   w[p + 'norm1.weight'], w[p + 'norm1.bias'])
    f = np.maximum(h @ w[p + 'linear1.weight'].T + w[p + 'linear1.bias'], 0.0)
    return layer_norm(h + f @ w[p + 'linear2.weight'].T + w[p + 'linear2.bias'], w[p + 'norm2.weight'], w[p + 'norm2.bias'])


def transition(prev_slots, weights, *, pred_type, num_layers=0, num_heads=0, norm_first=True, rnn=False,
               kernel_mlp=True, state=None, noise=None, dtype=np.float64):
    """One SAVi slot transition (reference savi.py:394-403): ``latents = predictor(prev_slots)`` (or, with
    ``pred_type=None``, ``latents = prev_slots``: the first frame's ``init_latents``), ``dist = kernel_dist_layer(latents)``,
    ``kernels = mu [+ noise * exp(log_var / 2)]``.

    ``weights``: StoSAVi state_dict entries (``predictor.*``, ``kernel_dist_layer.*``).  ``pred_type``: 'transformer'
    (predictor.py:20-44), 'mlp' (:47-74) or None; ``rnn``: wrapped in RNNPredictorWrapper (:76-113, one LSTM step over
    the B*K rows, ``state`` = (h, c) [B*K, H] or None for zeros).  Returns (dist [B,K,2D], kernels [B,K,D], new state)."""
    w = {k: np.asarray(v, dtype=dtype) for k, v in weights.items()}
    x = np.asarray(prev_slots, dtype=dtype)
    B, K, D = x.shape
    new_state = None
    if pred_type is not None:
        base = 'predictor.base_predictor.' if rnn else 'predictor.'
        if pred_type == 'transformer':
            for i in range(num_layers):
                p = f'{base}transformer_encoder.layers.{i}.'
                x = encoder_layer(x, w, p, num_heads, lambda a, bT: a @ bT.T) if norm_first \
                    else _post_ln_encoder_layer(x, w, p, num_heads)
        elif pred_type == 'mlp':
            normed = layer_norm(x, w[base + 'ln.weight'], w[base + 'ln.bias'])
            hid = np.maximum(normed @ w[base + 'mlp.0.weight'].T + w[base + 'mlp.0.bias'], 0.0)
            x = hid @ w[base + 'mlp.2.weight'].T + w[base + 'mlp.2.bias'] + (normed if norm_first else x)
        else:
            raise ValueError(pred_type)
        if rnn:
            H = w['predictor.rnn.weight_hh_l0'].shape[1]
            h, c = (np.zeros((B * K, H), dtype), np.zeros((B * K, H), dtype)) if state is None \
                else (np.asarray(state[0], dtype).reshape(B * K, H), np.asarray(state[1], dtype).reshape(B * K, H))
            g = x.reshape(B * K, D) @ w['predictor.rnn.weight_ih_l0'].T + w['predictor.rnn.bias_ih_l0'] \
                + h @ w['predictor.rnn.weight_hh_l0'].T + w['predictor.rnn.bias_hh_l0']
            i_, f_, g_, o_ = np.split(g, 4, axis=-1)                     # torch.nn.LSTM gate order
            c = _sigmoid(f_) * c + _sigmoid(i_) * np.tanh(g_)
            h = _sigmoid(o_) * np.tanh(c)
            new_state = (h, c)
            x = (h @ w['predictor.out_projector.weight'].T + w['predictor.out_projector.bias']).reshape(B, K, D)
    dist = x @ w['kernel_dist_layer.0.weight'].T + w['kernel_dist_layer.0.bias']
    if kernel_mlp:
        dist = np.maximum(layer_norm(dist, w['kernel_dist_layer.1.weight'], w['kernel_dist_layer.1.bias']), 0.0)
        dist = dist @ w['kernel_dist_layer.3.weight'].T + w['kernel_dist_layer.3.bias']
    kernels = dist[..., :D]
    if noise is not None:
        kernels = kernels + np.asarray(noise, dtype) * np.exp(0.5 * dist[..., D:])
    return dist, kernels, new_state


def decode_combine(dec_out, dtype=np.float64):
    """Tail of reference StoSAVi.decode (base_slots/models/savi.py:519-523).

    dec_out [B, K, 4, H, W] -> (recon_combined [B,3,H,W], masks [B,K,1,H,W]):
    masks = softmax over the slot axis of channel 3, recon_combined = sum_k dec_out[:, k, :3] * masks[:, k]."""
    x = np.asarray(dec_out, dtype=dtype)
    logit = x[:, :, 3:4]
    e = np.exp(logit - logit.max(axis=1, keepdims=True))
    masks = e / e.sum(axis=1, keepdims=True)
    return (x[:, :, :3] * masks).sum(axis=1), masks


def postproc_mask(batch_masks, fg_thre=0.5):
    """Reference postproc_mask (video_prediction/vp_utils.py:20-41), index arithmetic -> bit exact.

    batch_masks [B, T, N, 1, H, W] -> [B, T, H, W] int64.  np.argmin / np.argmax return the first
    extremum, as torch.argmin / torch.argmax do."""
    m = np.array(batch_masks, copy=True)
    B, T, N, _, H, W = m.shape
    m = m.reshape(B * T, N, H * W)
    bg_idx = m.max(-1).argmin(-1)                       # vp_utils.py:32-33
    bg_mask = m.max(1) < fg_thre                        # vp_utils.py:34-35
    for f in range(B * T):                              # vp_utils.py:36-39
        m[f, bg_idx[f], bg_mask[f]] = 1.
    return m.argmax(1).reshape(B, T, H, W).astype(np.int64)
