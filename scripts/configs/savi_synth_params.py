"""SAVi on synthetic OBJ3D-shaped clips (BASELINE config 1/2 shapes) -- written for this repo's smoke / DDP runs in the
reference's config convention (a `SlotFormerParams(BaseParams)` class loaded by file path, scripts/train.py:98-102)."""
from nerv.training import BaseParams


class SlotFormerParams(BaseParams):
    project = 'SlotFormer-B200'
    model = 'StoSAVi'
    gpus = 8
    max_epochs = 1
    lr = 1e-4
    clip_grad = 0.05
    warmup_steps_pct = 0.025
    train_batch_size = 32
    resolution = (64, 64)
    input_frames = 6
    slot_size = 128
    slot_dict = dict(num_slots=6, slot_size=slot_size, slot_mlp_size=slot_size * 2, num_iterations=2)
    enc_dict = dict(enc_channels=(3, 64, 64, 64, 64), enc_ks=5, enc_out_channels=slot_size, enc_norm='')
    dec_dict = dict(dec_channels=(slot_size, 64, 64, 64, 64), dec_resolution=(8, 8), dec_ks=5, dec_norm='')
    pred_dict = dict(pred_type='transformer', pred_rnn=True, pred_norm_first=True, pred_num_layers=2, pred_num_heads=4,
                     pred_ffn_dim=slot_size * 4, pred_sg_every=None)
    loss_dict = dict(use_post_recon_loss=True, kld_method='none')
    post_recon_loss_w = 1.
    kld_loss_w = 1.
