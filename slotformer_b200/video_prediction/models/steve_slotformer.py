"""SlotFormer over STEVE slots -- caller of hot path 2 for BASELINE config 4 (Physion).

Rollout side of reference ``STEVESlotFormer`` (slotformer/video_prediction/models/steve_slotformer.py:10-168):
the same ``SlotRollouter`` (the persistent sm_100a kernel) behind the same constructor and ``state_dict`` keys
(``rollouter.*``).  Its decoder -- the dVAE plus STEVE's autoregressive token decoder, "super slow" by the
reference's own words (steve_slotformer.py:87-90) -- is outside the hot-path scope (SURVEY.md section 2, rows 9,
12, 13) and is not constructed, so no dVAE / decoder checkpoint is needed; what it would feed (`decode`, the token
reconstruction loss) raises.  ``load_state_dict`` drops a released checkpoint's ``decoder.*`` / ``dvae.*`` entries.
"""
from torch.nn import functional as F

from .slotformer import SlotFormer

_SKIPPED_PREFIXES = ('decoder.', 'dvae.')


class STEVESlotFormer(SlotFormer):
    """Transformer-based rollouter on STEVE slot embeddings (rollout only)."""

    def __init__(self, resolution, clip_len,
                 slot_dict=dict(num_slots=6, slot_size=192),
                 dvae_dict=dict(down_factor=4, vocab_size=4096, dvae_ckp_path=''),
                 dec_dict=dict(dec_num_layers=4, dec_num_heads=4, dec_d_model=192, dec_ckp_path=''),
                 rollout_dict=dict(num_slots=6, slot_size=192, history_len=6, t_pe='sin', slots_pe='', d_model=192,
                                   num_layers=4, num_heads=8, ffn_dim=192 * 4, norm_first=True),
                 loss_dict=dict(rollout_len=6, use_img_recon_loss=False),
                 eps=1e-6):
        self.dvae_dict = dvae_dict
        super().__init__(resolution=resolution, clip_len=clip_len, slot_dict=slot_dict, dec_dict=dec_dict,
                         rollout_dict=rollout_dict, loss_dict=loss_dict, eps=eps)

    def _build_decoder(self):
        self.decoder = None                  # dVAE + token decoder: out of scope, see the module docstring
        self.decoder_pos_embedding = None

    def decode(self, slots):
        raise NotImplementedError('STEVESlotFormer.decode needs the dVAE / SLATE token decoder (out of scope)')

    def rollout(self, past_slots, pred_len, decode=False, with_gt=True):
        """steve_slotformer.py:105-109: always returns the predicted slots [B, pred_len, K, D]."""
        return self.rollouter(past_slots[:, -self.history_len:], pred_len)

    def forward(self, data_dict):
        slots = data_dict['slots']
        assert self.rollout_len + self.history_len == slots.shape[1], \
            f'wrong SlotFormer training length {slots.shape[1]}'
        if self.use_img_recon_loss:
            raise NotImplementedError('the token reconstruction loss needs the dVAE / SLATE token decoder (out of scope)')
        past, future = slots[:, :self.history_len], slots[:, self.history_len:]
        return {'gt_slots': future, 'pred_slots': self.rollout(past, self.rollout_len)}

    def calc_train_loss(self, data_dict, out_dict):
        return {'slot_recon_loss': F.mse_loss(out_dict['pred_slots'], out_dict['gt_slots'])}

    def train(self, mode=True):
        from ...compat.nerv.training import BaseModel
        return BaseModel.train(self, mode)

    def load_state_dict(self, state_dict, strict=True, **kw):
        kept = {k: v for k, v in state_dict.items() if not k.startswith(_SKIPPED_PREFIXES)}
        return super().load_state_dict(kept, strict=strict, **kw)
