"""Seeded construction of the wrapper models (StoSAVi / SlotFormer) for the wrapper goldens.
Weights come from torch's default initialisers under torch.manual_seed -- identical wherever the
same torch build runs (build container and GPU box share one image)."""
import numpy as np
import torch

SAVI_KW = dict(
    resolution=(64, 64), clip_len=3,
    slot_dict=dict(num_slots=5, slot_size=128, slot_mlp_size=256, num_iterations=2, kernel_mlp=True),
    enc_dict=dict(enc_channels=(3, 64, 64, 64, 64), enc_ks=5, enc_out_channels=128, enc_norm=''),
    dec_dict=dict(dec_channels=(128, 64, 64, 64, 64), dec_resolution=(8, 8), dec_ks=5, dec_norm=''),
    pred_dict=dict(pred_type='transformer', pred_rnn=True, pred_norm_first=True, pred_num_layers=2,
                   pred_num_heads=4, pred_ffn_dim=512, pred_sg_every=None),
    loss_dict=dict(use_post_recon_loss=True, kld_method='none'))

SLOTFORMER_KW = dict(
    resolution=(64, 64), clip_len=6,
    slot_dict=dict(num_slots=5, slot_size=128),
    dec_dict=dict(dec_channels=(128, 64, 64, 64, 64), dec_resolution=(8, 8), dec_ks=5, dec_norm='',
                  dec_ckp_path=''),
    rollout_dict=dict(num_slots=5, slot_size=128, history_len=6, t_pe='sin', slots_pe='',
                      d_model=128, num_layers=4, num_heads=8, ffn_dim=512, norm_first=True),
    loss_dict=dict(rollout_len=4, use_img_recon_loss=False))


def build_savi(cls):
    torch.manual_seed(0)
    return cls(**SAVI_KW).eval()


def build_slotformer(cls, ckpt_path):
    kw = dict(SLOTFORMER_KW)
    kw['dec_dict'] = dict(kw['dec_dict'], dec_ckp_path=ckpt_path)
    torch.manual_seed(1)
    return cls(**kw).eval()


def savi_input():
    rs = np.random.RandomState(7)
    return torch.from_numpy(rs.uniform(-1, 1, size=(2, 3, 3, 64, 64)).astype(np.float32))


def slotformer_input():
    rs = np.random.RandomState(8)
    return torch.from_numpy(rs.standard_normal((2, 10, 5, 128)).astype(np.float32))
