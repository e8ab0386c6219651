"""SAVi / StoSAVi video slot model -- the caller of hot path 1.

Same constructor, attributes, ``state_dict`` keys and methods as reference ``StoSAVi``
(slotformer/base_slots/models/savi.py:113-546): CNN encoder -> per-frame Slot Attention
(``self.slot_attention``, the sm_100a kernels) with slots carried through the predictor ->
optional spatial-broadcast decoder.  Encoder / predictor / decoder are stock PyTorch (cuDNN);
they are "next" rows f1-f3 of SURVEY.md section 8, not kernel targets of this round.
"""
import math

import torch
from torch import nn
from torch.nn import functional as F

from ...compat.nerv.models import conv_norm_act, deconv_norm_act, deconv_out_shape
from ...compat.nerv.training import BaseModel
from .predictor import ResidualMLPPredictor, RNNPredictorWrapper, TransformerPredictor
from ...engine import (ENC_TAIL_KEYS, RO_LAYER_KEYS, SFB_SA_NO_TCGEN05, SFB_TR_MLP, SFB_TR_TRANSFORMER, TR_MAX_LAYERS,
                       EncoderTailEngine, FeatureTiles, TransitionEngine)
from .slot_attention import SlotAttention
from .utils import SoftPositionEmbed, assert_shape, torch_cat


def build_broadcast_decoder(owner):
    """Creates ``owner.decoder`` / ``owner.decoder_pos_embedding`` from ``owner.dec_dict``
    (shared by StoSAVi and SlotFormer, reference savi.py:255-293)."""
    cfg = owner.dec_dict
    owner.dec_channels = cfg['dec_channels']
    owner.dec_resolution = cfg['dec_resolution']
    owner.dec_ks = cfg['dec_ks']
    owner.dec_norm = cfg['dec_norm']
    assert owner.dec_channels[0] == owner.slot_size, 'wrong in_channels for Decoder'
    blocks, size, stride = [], owner.dec_resolution[0], 2
    for cin, cout in zip(owner.dec_channels[:-1], owner.dec_channels[1:]):
        if size == owner.resolution[0]:
            stride = 1          # target size reached: remaining blocks keep the resolution
        blocks.append(deconv_norm_act(cin, cout, kernel_size=owner.dec_ks, stride=stride,
                                      norm=owner.dec_norm, act='relu'))
        size = deconv_out_shape(size, stride, owner.dec_ks // 2, owner.dec_ks, stride - 1)
    assert_shape(owner.resolution, (size, size),
                 message='Output shape of decoder did not match input resolution. '
                         'Try changing `decoder_resolution`.')
    blocks.append(nn.Conv2d(owner.dec_channels[-1], 4, kernel_size=1, stride=1, padding=0))
    owner.decoder = nn.Sequential(*blocks)
    owner.decoder_pos_embedding = SoftPositionEmbed(owner.slot_size, owner.dec_resolution)


def broadcast_decode(owner, slots):
    """slots [B, K, D] -> (recon_combined [B,3,H,W], recons [B,K,3,H,W], masks [B,K,1,H,W], slots).
    Spatial broadcast, deconv stack, softmax over slots of the alpha channel (savi.py:504-525)."""
    bs, num_slots, slot_size = slots.shape
    H, W = owner.resolution
    x = slots.reshape(bs * num_slots, slot_size, 1, 1).repeat(1, 1, *owner.dec_resolution)
    x = owner.decoder(owner.decoder_pos_embedding(x)).view(bs, num_slots, 4, H, W)
    if x.is_cuda and x.dtype == torch.float32 and not (torch.is_grad_enabled() and x.requires_grad) \
            and num_slots <= 12 and (H * W) % 4 == 0:
        # inference on the GPU: softmax over slots + recombination in one streaming kernel (csrc/decode_combine.cu)
        from ...engine import decode_combine
        recon_combined, masks = decode_combine(x)
        return recon_combined, x[:, :, :3], masks, slots
    recons, masks = x[:, :, :3], F.softmax(x[:, :, 3:], dim=1)
    return (recons * masks).sum(dim=1), recons, masks, slots


class StoSAVi(BaseModel):
    """Slot Attention for Video with an (optionally stochastic) slot-initialisation head.
    ``loss_dict['kld_method'] == 'none'`` gives plain deterministic SAVi."""

    def __init__(self, resolution, clip_len,
                 slot_dict=dict(num_slots=7, slot_size=128, slot_mlp_size=256, num_iterations=2,
                                kernel_mlp=True),
                 enc_dict=dict(enc_channels=(3, 64, 64, 64, 64), enc_ks=5, enc_out_channels=128,
                               enc_norm=''),
                 dec_dict=dict(dec_channels=(128, 64, 64, 64, 64), dec_resolution=(8, 8), dec_ks=5,
                               dec_norm=''),
                 pred_dict=dict(pred_type='transformer', pred_rnn=True, pred_norm_first=True,
                                pred_num_layers=2, pred_num_heads=4, pred_ffn_dim=512,
                                pred_sg_every=None),
                 loss_dict=dict(use_post_recon_loss=True, kld_method='var-0.01'),
                 eps=1e-6):
        super().__init__()
        self.resolution = resolution
        self.clip_len = clip_len
        self.eps = eps
        self.slot_dict, self.enc_dict, self.dec_dict = slot_dict, enc_dict, dec_dict
        self.pred_dict, self.loss_dict = pred_dict, loss_dict
        self._build_slot_attention()
        self._build_encoder()
        self._build_decoder()
        self._build_predictor()
        self._build_loss()
        self.testing = False        # True: forward returns slots only (offline extraction)

    # ------------------------------------------------------------------ build
    def _build_slot_attention(self):
        sd = self.slot_dict
        self.enc_out_channels = self.enc_dict['enc_out_channels']
        self.num_slots, self.slot_size = sd['num_slots'], sd['slot_size']
        self.slot_mlp_size, self.num_iterations = sd['slot_mlp_size'], sd['num_iterations']
        D = self.slot_size
        self.init_latents = nn.Parameter(nn.init.normal_(torch.empty(1, self.num_slots, D)))
        if sd.get('kernel_mlp', True):
            self.kernel_dist_layer = nn.Sequential(nn.Linear(D, 2 * D), nn.LayerNorm(2 * D),
                                                   nn.ReLU(), nn.Linear(2 * D, 2 * D))
        else:
            self.kernel_dist_layer = nn.Sequential(nn.Linear(D, 2 * D))
        # unused head kept so that released checkpoints load strictly
        self.prior_slot_layer = nn.Sequential(nn.Linear(D, D), nn.LayerNorm(D), nn.ReLU(),
                                              nn.Linear(D, D))
        self.slot_attention = SlotAttention(in_features=self.enc_out_channels,
                                            num_iterations=self.num_iterations,
                                            num_slots=self.num_slots, slot_size=D,
                                            mlp_hidden_size=self.slot_mlp_size, eps=self.eps)

    def _build_encoder(self):
        ed = self.enc_dict
        self.enc_channels = list(ed['enc_channels'])
        self.enc_ks, self.enc_norm = ed['enc_ks'], ed['enc_norm']
        self.visual_resolution = (64, 64)       # the feature grid is 64x64 for 64^2 and 128^2 input
        self.visual_channels = self.enc_channels[-1]
        n = len(self.enc_channels) - 1
        self.encoder = nn.Sequential(*[
            conv_norm_act(self.enc_channels[i], self.enc_channels[i + 1], kernel_size=self.enc_ks,
                          stride=2 if (i == 0 and self.resolution[0] == 128) else 1,
                          norm=self.enc_norm, act='relu' if i != n - 1 else '')
            for i in range(n)])
        self.encoder_pos_embedding = SoftPositionEmbed(self.visual_channels,
                                                       self.visual_resolution)
        self.encoder_out_layer = nn.Sequential(
            nn.LayerNorm(self.visual_channels),
            nn.Linear(self.visual_channels, self.enc_out_channels), nn.ReLU(),
            nn.Linear(self.enc_out_channels, self.enc_out_channels))

    def _build_decoder(self):
        build_broadcast_decoder(self)

    def _build_predictor(self):
        pd = self.pred_dict
        if pd.get('pred_type', 'transformer') == 'mlp':
            core = ResidualMLPPredictor([self.slot_size, self.slot_size * 2, self.slot_size],
                                        norm_first=pd['pred_norm_first'])
        else:
            core = TransformerPredictor(self.slot_size, pd['pred_num_layers'],
                                        pd['pred_num_heads'], pd['pred_ffn_dim'],
                                        norm_first=pd['pred_norm_first'])
        if pd['pred_rnn']:
            core = RNNPredictorWrapper(core, self.slot_size, self.slot_mlp_size, num_layers=1,
                                       rnn_cell='LSTM', sg_every=pd['pred_sg_every'])
        self.predictor = core

    def _build_loss(self):
        self.use_post_recon_loss = self.loss_dict['use_post_recon_loss']
        assert self.use_post_recon_loss
        method = self.loss_dict['kld_method']
        var = 1.0
        if '-' in method:
            method, v = method.split('-')
            var = float(v)
        self.kld_log_var = math.log(var)
        self.kld_method = method
        assert self.kld_method in ['var', 'none']

    # ------------------------------------------------------------------ pieces
    def _kld_loss(self, prior_dist, post_slots):
        """KL( N(mu, sigma) || N(mu, kld_var) ): only the variance is penalised."""
        if self.kld_method == 'none':
            return torch.tensor(0.).type_as(prior_dist)
        D = self.slot_size
        assert prior_dist.shape[-1] == 2 * D
        log_var1 = prior_dist[..., D:]
        log_var2 = torch.full_like(log_var1, self.kld_log_var)
        kld = 0.5 * (log_var2 - log_var1) + torch.exp(log_var1) / (2. * torch.exp(log_var2)) - 0.5
        return kld.sum(-1).mean()

    def _sample_dist(self, dist):
        D = self.slot_size
        assert dist.shape[-1] == 2 * D
        mu = dist[..., :D]
        if self.kld_method == 'none':
            return mu
        return mu + torch.randn_like(mu).detach() * torch.exp(0.5 * dist[..., D:])

    def _get_encoder_out(self, img):
        """img [N, 3, H, W] -> [N, 4096, enc_out_channels] feature vectors."""
        x = self.encoder(img).type(self.dtype)
        x = self.encoder_pos_embedding(x).flatten(2, 3).permute(0, 2, 1).contiguous()
        return self.encoder_out_layer(x)

    # -- encoder tail fused into the Slot Attention input format ------------------------------
    # In inference the MLP after the CNN (+ positional embedding, both LayerNorms) runs as ONE sm_100a kernel that
    # writes the fp16 operand tiles Slot Attention streams (csrc/enc_tail.cu): the fp32 [B*T, 4096, C] feature grid
    # of the reference is never formed.  Shapes outside the kernel's envelope keep the stock layers.
    fuse_encoder_tail = True

    def _encoder_tail_fusable(self, img):
        sa = self.slot_attention
        return (self.fuse_encoder_tail and img.is_cuda and not self.training and not torch.is_grad_enabled()
                and self.enc_out_channels == 128 and self.visual_channels == 64 and self.dtype == torch.float32
                and sa.in_features == 128 and sa.slot_size == 128 and sa.num_slots <= 8
                and not (sa.engine_flags & SFB_SA_NO_TCGEN05))

    def _get_encoder_tiles(self, img):
        """img [N, 3, H, W] -> FeatureTiles [N] (same role as _get_encoder_out, savi.py:367-377)."""
        if '_enc_tail_engine' not in self.__dict__:
            self.__dict__['_enc_tail_engine'] = EncoderTailEngine()
        named = dict(self.named_parameters())
        # the cuDNN CNN runs channels-last (the native layout of its tensor-core kernels: 3.7 ms instead of 5.1 ms for
        # 384 frames of 128 x 128; same TF32 arithmetic); the tail kernel reads that layout directly (SFB_ET_NHWC)
        if not self.__dict__.get('_encoder_channels_last'):
            self.encoder.to(memory_format=torch.channels_last)
            self.__dict__['_encoder_channels_last'] = True
        x = self.encoder(img.contiguous(memory_format=torch.channels_last)).type(self.dtype)
        return self._enc_tail_engine.forward(x, {k: named[k].detach() for k in ENC_TAIL_KEYS}, self.enc_out_channels,
                                             max_ctas=self.slot_attention.max_ctas)

    # -- predictor -> kernel_dist_layer -> sample as ONE kernel per frame -------------------
    # The slot transition between two Slot Attention calls (reference savi.py:393-410 with predictor.py:20-113) is
    # ~40 stock launches on [B*K, D] activations; in inference it runs as one cluster-per-clip kernel
    # (csrc/transition.cu).  Structures outside the kernel's envelope keep the stock modules.
    fuse_transition = True

    def _transition_spec(self):
        """Structure + weights of ``predictor`` / ``kernel_dist_layer`` for sfb_transition_forward, or None."""
        D = self.slot_size
        W = {}
        spec = dict(pred_type=0, num_layers=0, num_heads=0, ffn_dim=0, norm_first=True, mlp_hidden=0, rnn_hidden=0,
                    weights=W)

        def is_ln(m, width):
            return isinstance(m, nn.LayerNorm) and m.elementwise_affine and m.bias is not None and m.eps == 1e-5 \
                and tuple(m.normalized_shape) == (width,)

        def is_lin(m, cin, cout):
            return isinstance(m, nn.Linear) and m.bias is not None and m.in_features == cin and m.out_features == cout

        core = self.predictor
        if isinstance(core, RNNPredictorWrapper):
            rnn = core.rnn
            if not isinstance(rnn, nn.LSTM) or rnn.num_layers != 1 or rnn.bidirectional or not rnn.bias \
                    or rnn.proj_size != 0 or rnn.input_size != D or not is_lin(core.out_projector, rnn.hidden_size, D):
                return None
            spec['rnn_hidden'] = rnn.hidden_size
            for k in ('weight_ih_l0', 'weight_hh_l0', 'bias_ih_l0', 'bias_hh_l0'):
                W['rnn.' + k] = getattr(rnn, k).detach()
            W['out_projector.weight'] = core.out_projector.weight.detach()
            W['out_projector.bias'] = core.out_projector.bias.detach()
            core = core.base_predictor
        if isinstance(core, TransformerPredictor):
            enc = core.transformer_encoder
            layers = list(enc.layers)
            l0 = layers[0]
            if enc.norm is not None or not 1 <= len(layers) <= TR_MAX_LAYERS:
                return None
            for i, ly in enumerate(layers):
                att = ly.self_attn
                relu = ly.activation is F.relu or getattr(ly, 'activation_relu_or_gelu', 0) == 1
                if not relu or att.embed_dim != D or att.in_proj_weight is None or att.in_proj_bias is None \
                        or att.bias_k is not None or att.add_zero_attn or ly.norm_first != l0.norm_first \
                        or att.num_heads != l0.self_attn.num_heads or not is_ln(ly.norm1, D) or not is_ln(ly.norm2, D) \
                        or not is_lin(ly.linear1, D, l0.linear1.out_features) or not is_lin(ly.linear2, l0.linear1.out_features, D):
                    return None
                for k in RO_LAYER_KEYS:
                    obj = ly
                    for part in k.split('.'):
                        obj = getattr(obj, part)
                    W[f'layers.{i}.{k}'] = obj.detach()
            spec.update(pred_type=SFB_TR_TRANSFORMER, num_layers=len(layers), num_heads=l0.self_attn.num_heads,
                        ffn_dim=l0.linear1.out_features, norm_first=bool(l0.norm_first))
        elif isinstance(core, ResidualMLPPredictor):
            mlp = core.mlp
            if len(mlp) != 3 or not isinstance(mlp[1], nn.ReLU) or not is_ln(core.ln, D) \
                    or not isinstance(mlp[0], nn.Linear) or not is_lin(mlp[0], D, mlp[0].out_features) \
                    or not is_lin(mlp[2], mlp[0].out_features, D):
                return None
            spec.update(pred_type=SFB_TR_MLP, mlp_hidden=mlp[0].out_features, norm_first=bool(core.norm_first))
            for name, m in (('ln', core.ln), ('mlp.0', mlp[0]), ('mlp.2', mlp[2])):
                W[name + '.weight'] = m.weight.detach()
                W[name + '.bias'] = m.bias.detach()
        else:
            return None
        kd = self.kernel_dist_layer
        if len(kd) == 4 and is_lin(kd[0], D, 2 * D) and is_ln(kd[1], 2 * D) and isinstance(kd[2], nn.ReLU) \
                and is_lin(kd[3], 2 * D, 2 * D):
            spec['kernel_mlp'] = True
            mods = (('0', kd[0]), ('1', kd[1]), ('3', kd[3]))
        elif len(kd) == 1 and is_lin(kd[0], D, 2 * D):
            spec['kernel_mlp'] = False
            mods = (('0', kd[0]),)
        else:
            return None
        for name, m in mods:
            W[f'kernel_dist.{name}.weight'] = m.weight.detach()
            W[f'kernel_dist.{name}.bias'] = m.bias.detach()
        return spec

    def _transition_fusable(self, feats):
        if not (self.fuse_transition and feats.is_cuda and not self.training and not torch.is_grad_enabled()
                and self.dtype == torch.float32 and self.num_slots <= 8):
            return None
        # one cluster (>= 1 CTA) per clip, a single wave: beyond that many clips the stock GEMMs on [B*K, D] win
        if feats.shape[0] > torch.cuda.get_device_properties(feats.device).multi_processor_count:
            return None
        spec = self._transition_spec()
        if spec is None or any(not t.is_cuda or t.dtype != torch.float32 for t in spec['weights'].values()) \
                or not TransitionEngine.supported(spec, self.slot_size):
            return None
        return spec

    def _transition(self, spec, prev_slots, B):
        """(kernel_dist, initial slots) of the next frame from the previous frame's slots: one kernel launch."""
        if '_transition_engine' not in self.__dict__:
            self.__dict__['_transition_engine'] = TransitionEngine()
        pred = self.predictor
        rnn = spec['rnn_hidden'] > 0 and prev_slots is not None
        state = pred.hidden_state if rnn else None
        noise = None
        if self.kld_method != 'none':
            noise = torch.randn((B, self.num_slots, self.slot_size), dtype=torch.float32, device=self.init_latents.device)
        prev = self.init_latents.detach() if prev_slots is None else prev_slots
        dist, init, new = self._transition_engine.forward(spec, prev, prev_slots is not None, B, state, noise)
        if rnn:
            pred.hidden_state = (new[0].unsqueeze(0), new[1].unsqueeze(0))
            pred.step += 1
        return dist, init

    def _frame_loop(self, feats, prev_slots):
        """The serial per-frame chain (reference savi.py:393-410): predictor -> kernel_dist_layer -> sample ->
        Slot Attention, T times.  feats [B, T, N, C] -> (kernel_dist [B,T,K,2D], post_slots [B,T,K,D])."""
        B, T = feats.shape[:2]
        spec = self._transition_fusable(feats)
        start = self.init_latents.repeat(B, 1, 1) if spec is None else None
        dists, slots = [], []
        for t in range(T):                                  # frames are a serial chain
            if spec is not None:
                dist, init = self._transition(spec, prev_slots, B)
            else:
                latents = start if prev_slots is None else self.predictor(prev_slots)
                dist = self.kernel_dist_layer(latents)
                init = self._sample_dist(dist)
            prev_slots = self.slot_attention(feats[:, t], init)   # hot path 1
            dists.append(dist)
            slots.append(prev_slots)
        return torch.stack(dists, dim=1), torch.stack(slots, dim=1)

    # -- the frame loop as ONE CUDA-graph launch -------------------------------------
    # Per frame the loop issues ~45 small launches (2-layer Transformer + LSTM predictor, the distribution head
    # and the 5 Slot Attention kernels); at SAVi batch sizes that is launch-bound.  In inference the whole T-frame
    # chain is captured once per (B, T, first-frame mode) and replayed: same kernels, same order, same results.
    use_cuda_graph = True

    def _graph_key(self, feats, prev_slots, has_state):
        ver = tuple((p.data_ptr(), p._version) for p in self.parameters())
        return (tuple(feats.shape), feats.dtype, feats.device, prev_slots is not None, has_state, ver)

    @staticmethod
    def _state_tensors(state):
        if state is None:
            return None
        return (state,) if isinstance(state, torch.Tensor) else tuple(state)

    def _frame_loop_graphed(self, feats, prev_slots):
        # the recurrent state of the predictor is Python-side module state: it enters the captured loop through
        # static buffers and leaves it through the captured final-state tensors
        pred = self.predictor
        stateful = hasattr(pred, 'hidden_state')
        h_in = self._state_tensors(pred.hidden_state) if stateful else None
        single = stateful and isinstance(pred.hidden_state, torch.Tensor)
        if hasattr(pred, 'rnn'):
            # before the key: flattening re-points the weights.  nn.RNNBase.flatten_parameters() allocates a NEW flat
            # buffer on every call (the parameters then alternate between two addresses and every other call would
            # re-capture the graph and re-pack the weights): flatten once per set of weight addresses
            ptrs = tuple(w.data_ptr() for w in pred.rnn._flat_weights)
            if self.__dict__.get('_rnn_flat_ptrs') != ptrs:
                pred.rnn.flatten_parameters()
                self.__dict__['_rnn_flat_ptrs'] = tuple(w.data_ptr() for w in pred.rnn._flat_weights)
        key = self._graph_key(feats, prev_slots, h_in is not None)
        tiled = isinstance(feats, FeatureTiles)
        raw = feats.data if tiled else feats
        cache = self.__dict__.setdefault('_loop_graphs', {})
        ent = cache.get(key)
        if ent is None:
            if len(cache) >= 4:
                cache.clear()                               # weights changed or many shapes: start over
            s_raw = torch.empty_like(raw)
            s_feats = FeatureTiles(s_raw, feats.N, feats.C) if tiled else s_raw
            s_prev = torch.empty_like(prev_slots) if prev_slots is not None else None
            s_h = tuple(torch.empty_like(h) for h in h_in) if h_in is not None else None
            s_raw.copy_(raw)
            if s_prev is not None:
                s_prev.copy_(prev_slots)
            if s_h is not None:
                for d, h in zip(s_h, h_in):
                    d.copy_(h)
            step0 = pred.step if stateful else 0

            def run():
                if stateful:
                    pred.hidden_state = None if s_h is None else (s_h[0] if single else s_h)
                    pred.step = step0
                out = self._frame_loop(s_feats, s_prev)
                return out, (self._state_tensors(pred.hidden_state) if stateful else None)

            side = torch.cuda.Stream(feats.device)
            side.wait_stream(torch.cuda.current_stream(feats.device))
            with torch.cuda.stream(side):                   # warm-up outside the capture (lazy handles, workspaces)
                run()
            torch.cuda.current_stream(feats.device).wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                out, h_out = run()
            ent = cache[key] = (graph, s_raw, s_prev, s_h, out, h_out, (pred.step - step0) if stateful else 0)
            if stateful:
                pred.step = step0
        graph, s_raw, s_prev, s_h, out, h_out, nsteps = ent
        s_raw.copy_(raw)
        if s_prev is not None:
            s_prev.copy_(prev_slots)
        if s_h is not None:
            for d, h in zip(s_h, h_in):
                d.copy_(h)
        graph.replay()
        if stateful:
            if h_out is None:
                pred.hidden_state = None
            else:
                new = tuple(h.clone() for h in h_out)
                pred.hidden_state = new[0] if (len(new) == 1 and not isinstance(self.predictor.rnn, nn.LSTM)) else new
            pred.step += nsteps
        return out[0].clone(), out[1].clone()

    def encode_features(self, img):
        """img [B, T, 3, H, W] -> per-pixel features [B, T, N, C] (the first half of ``encode``: CNN + encoder tail; in
        inference the fp16 operand tiles of the fused tail).  ``encode(img, feats=...)`` takes them back, so a caller
        can run this stage of clip batch i+1 on another stream while the serial frame loop of batch i is in flight."""
        B, T = img.shape[:2]
        if self._encoder_tail_fusable(img):
            return self._get_encoder_tiles(img.flatten(0, 1)).unflatten(0, (B, T))
        return self._get_encoder_out(img.flatten(0, 1)).unflatten(0, (B, T))

    def encode(self, img, prev_slots=None, feats=None):
        """img [B, T, 3, H, W] -> (kernel_dist [B,T,K,2D], post_slots [B,T,K,D], features).
        ``feats``: the result of ``encode_features(img)`` if the caller computed it already."""
        if feats is None:
            feats = self.encode_features(img)
        if self.use_cuda_graph and feats.is_cuda and not torch.is_grad_enabled() and not self.training:
            try:
                dists, slots = self._frame_loop_graphed(feats.contiguous(), prev_slots)
                return dists, slots, feats
            except RuntimeError as e:                       # an op that cannot be captured: stay eager from now on
                if 'captur' not in str(e).lower() and 'graph' not in str(e).lower():
                    raise
                self.use_cuda_graph = False
        dists, slots = self._frame_loop(feats, prev_slots)
        return dists, slots, feats

    def _reset_rnn(self):
        self.predictor.reset()

    def decode(self, slots):
        return broadcast_decode(self, slots)

    # ------------------------------------------------------------------ forward
    def forward(self, data_dict):
        """Long evaluation videos are cut into temporal chunks (halving the chunk until it fits
        in memory) with the slots carried across chunks."""
        img = data_dict['img']
        T = img.shape[1]
        if T <= self.clip_len or self.training:
            return self._forward(img, None)
        clip_len = T
        while True:
            try:
                probe = self._forward(img[:, :clip_len], None)
                del probe
                torch.cuda.empty_cache()
                break
            except RuntimeError:                            # CUDA out of memory
                clip_len = clip_len // 2 + 1
        self.clip_len = max(self.clip_len, clip_len)
        if clip_len == T:
            return self._forward(img, None)
        pieces, prev_slots = None, None
        for t0 in range(0, T, clip_len):
            out = self._forward(img[:, t0:t0 + clip_len], prev_slots)
            if pieces is None:
                pieces = {k: [v.detach()] for k, v in out.items()}
            else:
                for k, v in out.items():
                    pieces[k].append(v.detach())
            prev_slots = pieces['post_slots'][-1][:, -1].detach().clone()
            del out
            torch.cuda.empty_cache()
        return {k: torch_cat(v, dim=1) for k, v in pieces.items()}

    def _forward(self, img, prev_slots=None):
        if prev_slots is None:
            self._reset_rnn()
        B, T = img.shape[:2]
        kernel_dist, post_slots, _ = self.encode(img, prev_slots=prev_slots)
        out = {'post_slots': post_slots, 'kernel_dist': kernel_dist, 'img': img}
        if self.testing:
            return out
        if self.use_post_recon_loss:
            combined, recons, masks, _ = self.decode(post_slots.flatten(0, 1))
            out.update(post_recon_combined=combined.unflatten(0, (B, T)),
                       post_recons=recons.unflatten(0, (B, T)),
                       post_masks=masks.unflatten(0, (B, T)))
        return out

    def calc_train_loss(self, data_dict, out_dict):
        losses = {'kld_loss': self._kld_loss(out_dict['kernel_dist'], out_dict['post_slots'])}
        if self.use_post_recon_loss:
            losses['post_recon_loss'] = F.mse_loss(out_dict['post_recon_combined'],
                                                   out_dict['img'])
        return losses

    @property
    def dtype(self):
        return self.slot_attention.dtype

    @property
    def device(self):
        return self.slot_attention.device
