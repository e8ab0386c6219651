"""STEVE slot extraction -- the second caller of hot path 1 (BASELINE config 4, Physion).

Restates the slot side of reference ``STEVE`` (slotformer/base_slots/models/steve.py:76-350): CNN encoder ->
per-frame ``SlotAttentionWMask`` (the sm_100a kernels, seg mask included) with the slots carried through the
Transformer + LSTM predictor -> bilinear up-sampling of the masks in eval (steve.py:198-240).  Same constructor
arguments and the same ``state_dict`` keys for everything it builds.

Out of scope: the dVAE tokenizer and the SLATE token decoder, i.e. STEVE's
training loss.  They are not constructed: no dVAE checkpoint is needed, ``forward`` serves the extraction path
(``testing = True``, what extract_slots.py runs: steve.py:305-307), and asking for the token losses raises.
A released STEVE checkpoint loads with ``load_state_dict``: its ``dvae.*`` / ``trans_decoder.*`` entries are
dropped, everything else must match strictly.
"""
import torch
from torch import nn
from torch.nn import functional as F

from ...compat.nerv.training import BaseModel
from .savi import StoSAVi
from .slot_attention import SlotAttentionWMask

_SKIPPED_PREFIXES = ('dvae.', 'trans_decoder.')


class STEVE(StoSAVi):
    """Slot-extraction half of STEVE (Slot Attention with segmentation masks over video)."""

    def __init__(self, resolution, clip_len,
                 slot_dict=dict(num_slots=7, slot_size=128, slot_mlp_size=256, num_iterations=2),
                 dvae_dict=dict(down_factor=4, vocab_size=4096, dvae_ckp_path=''),
                 enc_dict=dict(enc_channels=(3, 64, 64, 64, 64), enc_ks=5, enc_out_channels=128, enc_norm=''),
                 dec_dict=dict(dec_type='slate', dec_num_layers=4, dec_num_heads=4, dec_d_model=128),
                 pred_dict=dict(pred_rnn=True, pred_norm_first=True, pred_num_layers=2, pred_num_heads=4,
                                pred_ffn_dim=512, pred_sg_every=None),
                 loss_dict=dict(use_img_recon_loss=False),
                 eps=1e-6):
        BaseModel.__init__(self)
        self.resolution = resolution
        self.clip_len = clip_len
        self.eps = eps
        self.slot_dict, self.dvae_dict, self.enc_dict = slot_dict, dvae_dict, enc_dict
        self.dec_dict, self.pred_dict, self.loss_dict = dec_dict, pred_dict, loss_dict
        self._build_slot_attention()
        self._build_encoder()
        self._build_predictor()
        self._build_loss()
        self.testing = True         # only the extraction path exists here

    def _build_slot_attention(self):
        sd = self.slot_dict
        self.enc_out_channels = self.enc_dict['enc_out_channels']
        self.num_slots, self.slot_size = sd['num_slots'], sd['slot_size']
        self.slot_mlp_size, self.num_iterations = sd['slot_mlp_size'], sd['num_iterations']
        # learnable initial slots, used directly (no kernel distribution head in STEVE)
        self.init_latents = nn.Parameter(nn.init.normal_(torch.empty(1, self.num_slots, self.slot_size)))
        self.slot_attention = SlotAttentionWMask(in_features=self.enc_out_channels,
                                                 num_iterations=self.num_iterations, num_slots=self.num_slots,
                                                 slot_size=self.slot_size, mlp_hidden_size=self.slot_mlp_size,
                                                 eps=self.eps)

    def _build_loss(self):
        self.use_img_recon_loss = self.loss_dict['use_img_recon_loss']

    def encode(self, img, prev_slots=None):
        """img [B, T, 3, H, W] -> (slots [B,T,K,D], masks [B,T,K,H,W], features) -- steve.py:198-240."""
        B, T = img.shape[:2]
        feats = self._get_encoder_out(img.flatten(0, 1)).unflatten(0, (B, T))
        all_slots, all_masks = [], []
        for t in range(T):                                   # frames are a serial chain
            latents = self.init_latents.repeat(B, 1, 1) if prev_slots is None else self.predictor(prev_slots)
            prev_slots, masks = self.slot_attention(feats[:, t], latents)      # hot path 1 (mask variant)
            all_slots.append(prev_slots)
            all_masks.append(masks.unflatten(-1, self.visual_resolution))
        slots = torch.stack(all_slots, dim=1)
        masks = torch.stack(all_masks, dim=1).contiguous()
        if not self.training and tuple(self.visual_resolution) != tuple(self.resolution):
            with torch.no_grad():
                masks = F.interpolate(masks.flatten(0, 2).unsqueeze(1), self.resolution, mode='bilinear',
                                      align_corners=False).squeeze(1).unflatten(0, (B, T, self.num_slots))
        return slots, masks, feats

    def _forward(self, img, prev_slots=None):
        if prev_slots is None:
            self._reset_rnn()
        slots, masks, _ = self.encode(img, prev_slots)
        if not self.testing:
            raise NotImplementedError('STEVE token losses need the dVAE tokenizer and the SLATE decoder, which are '
                                      'outside the hot-path scope (SURVEY.md section 2, rows 12-13)')
        # 'post_slots' duplicates 'slots' so that StoSAVi.forward's temporal chunking (which carries post_slots) applies
        return {'slots': slots, 'masks': masks, 'post_slots': slots}

    def forward(self, data_dict):
        out = super().forward(data_dict)
        out.pop('post_slots', None)
        return out

    def decode(self, slots):
        raise NotImplementedError('the SLATE token decoder is outside the hot-path scope')

    def calc_train_loss(self, data_dict, out_dict):
        raise NotImplementedError('STEVE training (token losses) is outside the hot-path scope')

    def load_state_dict(self, state_dict, strict=True, **kw):
        kept = {k: v for k, v in state_dict.items() if not k.startswith(_SKIPPED_PREFIXES)}
        return super().load_state_dict(kept, strict=strict, **kw)
