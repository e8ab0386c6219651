"""SlotFormer dynamics model -- the caller of hot path 2.

Same constructor, attributes, ``state_dict`` keys and methods as reference ``SlotFormer``
(slotformer/video_prediction/models/slotformer.py:137-343): a ``SlotRollouter`` (the persistent
sm_100a rollout kernel) plus the frozen SAVi spatial-broadcast decoder loaded from a SAVi
checkpoint (needed only for image reconstruction losses / visualisation)."""
import torch
from torch.nn import functional as F

from ...base_slots.models.savi import broadcast_decode, build_broadcast_decoder
from ...compat.nerv.training import BaseModel
from .rollouter import SlotRollouter, get_sin_pos_enc, build_pos_enc  # noqa: F401


class SlotFormer(BaseModel):
    """Transformer-based autoregressive dynamics model over slots."""

    def __init__(self, resolution, clip_len,
                 slot_dict=dict(num_slots=7, slot_size=128),
                 dec_dict=dict(dec_channels=(128, 64, 64, 64, 64), dec_resolution=(8, 8), dec_ks=5,
                               dec_norm='', dec_ckp_path=''),
                 rollout_dict=dict(num_slots=7, slot_size=128, history_len=6, t_pe='sin',
                                   slots_pe='', d_model=128, num_layers=4, num_heads=8,
                                   ffn_dim=512, norm_first=True),
                 loss_dict=dict(rollout_len=6, use_img_recon_loss=False),
                 eps=1e-6):
        super().__init__()
        self.resolution = resolution
        self.clip_len = clip_len
        self.eps = eps
        self.slot_dict, self.dec_dict = slot_dict, dec_dict
        self.rollout_dict, self.loss_dict = rollout_dict, loss_dict
        self._build_slot_attention()
        self._build_decoder()
        self._build_rollouter()
        self._build_loss()
        self.testing = False
        self.loss_decay_factor = 1.         # temporal loss weighting, set by the trainer

    def _build_slot_attention(self):
        self.num_slots = self.slot_dict['num_slots']
        self.slot_size = self.slot_dict['slot_size']

    def _build_decoder(self):
        """Decoder architecture of SAVi; weights come from ``dec_ckp_path`` and stay frozen."""
        build_broadcast_decoder(self)
        path = self.dec_dict['dec_ckp_path']
        assert path, 'Please provide pretrained decoder weight'
        state = torch.load(path, map_location='cpu')['state_dict']

        def strip(prefix):
            return {k[len(prefix):]: v for k, v in state.items() if k.startswith(prefix)}

        self.decoder.load_state_dict(strip('decoder.'))
        self.decoder_pos_embedding.load_state_dict(strip('decoder_pos_embedding.'))
        for module in (self.decoder, self.decoder_pos_embedding):
            for p in module.parameters():
                p.requires_grad = False
            module.eval()

    def _build_rollouter(self):
        self.history_len = self.rollout_dict['history_len']
        self.rollouter = SlotRollouter(**self.rollout_dict)

    def _build_loss(self):
        self.rollout_len = self.loss_dict['rollout_len']
        self.use_img_recon_loss = self.loss_dict['use_img_recon_loss']

    def decode(self, slots):
        return broadcast_decode(self, slots)

    def rollout(self, past_slots, pred_len, decode=False, with_gt=True):
        """past_slots [B, T, K, D]; predicts ``pred_len`` future steps from the last
        ``history_len`` frames (hot path 2); optionally decodes images."""
        B = past_slots.shape[0]
        pred_slots = self.rollouter(past_slots[:, -self.history_len:], pred_len)
        if not decode:
            return pred_slots
        slots = torch.cat([past_slots, pred_slots], dim=1) if with_gt else pred_slots
        T = slots.shape[1]
        combined, recons, masks, _ = self.decode(slots.flatten(0, 1))
        return {'recon_combined': combined.unflatten(0, (B, T)),
                'recons': recons.unflatten(0, (B, T)),
                'masks': masks.unflatten(0, (B, T)),
                'slots': slots}

    def forward(self, data_dict):
        slots = data_dict['slots']
        assert self.rollout_len + self.history_len == slots.shape[1], \
            f'wrong SlotFormer training length {slots.shape[1]}'
        past, future = slots[:, :self.history_len], slots[:, self.history_len:]
        if self.use_img_recon_loss:
            out = self.rollout(past, self.rollout_len, decode=True, with_gt=False)
            out['pred_slots'] = out.pop('slots')
            out['gt_slots'] = future
            return out
        return {'gt_slots': future, 'pred_slots': self.rollout(past, self.rollout_len)}

    def calc_train_loss(self, data_dict, out_dict):
        gt, pred = out_dict['gt_slots'], out_dict['pred_slots']
        err = F.mse_loss(pred, gt, reduction='none')            # [B, T, K, D]
        losses = {}
        if not self.training:
            for step in range(min(6, gt.shape[1])):
                losses[f'slot_recon_loss_{step + 1}'] = err[:, step].mean()
        if self.loss_decay_factor < 1.:
            w = (self.loss_decay_factor ** torch.arange(gt.shape[1])).type_as(err)
            err = err * (w / w.sum() * gt.shape[1])[None, :, None, None]
        vid_len = data_dict.get('vid_len', None)
        horizon = self.history_len + self.rollout_len
        valid = None
        if vid_len is not None and (vid_len < horizon).any():   # PHYRE: videos of different length
            t = torch.arange(gt.shape[1], device=gt.device) + self.history_len
            valid = (t[None] < vid_len[:, None]).flatten(0, 1)
            err = err.flatten(0, 1)[valid]
        losses['slot_recon_loss'] = err.mean()
        if self.use_img_recon_loss:
            img_err = F.mse_loss(out_dict['recon_combined'], data_dict['img'][:, self.history_len:],
                                 reduction='none')
            if valid is not None:
                img_err = img_err.flatten(0, 1)[valid]
            losses['img_recon_loss'] = img_err.mean()
        return losses

    @property
    def dtype(self):
        return self.rollouter.dtype

    @property
    def device(self):
        return self.rollouter.device

    def train(self, mode=True):
        super().train(mode)
        self.decoder.eval()                 # the decoder is frozen
        self.decoder_pos_embedding.eval()
        return self
