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


# ---- STEVE / STEVESlotFormer (BASELINE config 4 callers): D = 192 slots, 128x128 frames, masks up-sampled ----
STEVE_KW = dict(
    resolution=(128, 128), clip_len=3,
    slot_dict=dict(num_slots=6, slot_size=192, slot_mlp_size=384, num_iterations=2),
    dvae_dict=dict(down_factor=4, vocab_size=64, dvae_ckp_path=''),
    enc_dict=dict(enc_channels=(3, 64, 64, 64, 64), enc_ks=5, enc_out_channels=192, enc_norm=''),
    dec_dict=dict(dec_type='slate', dec_num_layers=1, dec_num_heads=4, dec_d_model=192),
    pred_dict=dict(pred_rnn=True, pred_norm_first=True, pred_num_layers=2, pred_num_heads=4, pred_ffn_dim=512,
                   pred_sg_every=None),
    loss_dict=dict(use_img_recon_loss=False))

STEVE_SLOTFORMER_KW = dict(
    resolution=(128, 128), clip_len=6,
    slot_dict=dict(num_slots=6, slot_size=192),
    dvae_dict=dict(down_factor=4, vocab_size=64, dvae_ckp_path=''),
    dec_dict=dict(dec_num_layers=1, dec_num_heads=4, dec_d_model=192, dec_ckp_path=''),
    rollout_dict=dict(num_slots=6, slot_size=192, history_len=6, t_pe='sin', slots_pe='', d_model=256, num_layers=2,
                      num_heads=8, ffn_dim=1024, norm_first=True),
    loss_dict=dict(rollout_len=4, use_img_recon_loss=False))


def fill_seeded(model, seed, skip=()):
    """Deterministic weights that do not depend on construction order: every state_dict entry (sorted by key,
    `skip` prefixes left alone) is drawn from a numpy stream -- matrices ~ N(0, 1/fan_in), vectors ~ 0.05 N(0, 1),
    LayerNorm gains 1 + 0.1 N(0, 1).  The same call on the reference model and on ours gives identical weights."""
    rs = np.random.RandomState(seed)
    sd = model.state_dict()
    with torch.no_grad():
        for k in sorted(sd):
            if k.startswith(tuple(skip)) or k.endswith('enc_t_pe'):
                continue
            v = sd[k]
            x = rs.standard_normal(tuple(v.shape)).astype(np.float32)
            if v.dim() >= 2:
                x /= np.sqrt(float(np.prod(v.shape[1:])))
            elif 'norm' in k.lower() and k.endswith('weight') or k.endswith(('.0.weight',)) and v.dim() == 1:
                x = 1. + 0.1 * x
            else:
                x *= 0.05
            v.copy_(torch.from_numpy(x))
    return model


def build_steve(cls, **override):
    kw = dict(STEVE_KW)
    kw.update(override)
    torch.manual_seed(2)
    return fill_seeded(cls(**kw).eval(), 102, skip=('dvae.', 'trans_decoder.'))


def build_steve_slotformer(cls, **override):
    kw = dict(STEVE_SLOTFORMER_KW)
    kw.update(override)
    torch.manual_seed(3)
    return fill_seeded(cls(**kw).eval(), 103, skip=('dvae.', 'decoder.'))


def steve_input():
    rs = np.random.RandomState(9)
    return torch.from_numpy(rs.uniform(-1, 1, size=(2, 3, 3, 128, 128)).astype(np.float32))


def steve_slotformer_input():
    rs = np.random.RandomState(10)
    return torch.from_numpy(rs.standard_normal((2, 10, 6, 192)).astype(np.float32))
