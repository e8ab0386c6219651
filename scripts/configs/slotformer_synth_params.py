"""SlotFormer on synthetic OBJ3D-shaped slot sequences (BASELINE config 2: K = 6, T 6 -> 10, d = 128, 4 layers) in the
reference's config convention.  `dec_ckp_path` is empty: scripts/train.py --synthetic-decoder mints a decoder
checkpoint from a freshly initialised SAVi (the reference needs a trained one, slotformer.py:201-210)."""
from nerv.training import BaseParams


class SlotFormerParams(BaseParams):
    project = 'SlotFormer-B200'
    model = 'SlotFormer'
    gpus = 8
    max_epochs = 1
    lr = 2e-4
    clip_grad = -1.
    warmup_steps_pct = 0.05
    train_batch_size = 512
    resolution = (64, 64)
    input_frames = 6
    frame_offset = 1
    slot_size = 128
    slot_dict = dict(num_slots=6, slot_size=slot_size)
    dec_dict = dict(dec_channels=(slot_size, 64, 64, 64, 64), dec_resolution=(8, 8), dec_ks=5, dec_norm='', dec_ckp_path='')
    rollout_dict = dict(num_slots=6, slot_size=slot_size, history_len=input_frames, t_pe='sin', slots_pe='', d_model=128,
                        num_layers=4, num_heads=8, ffn_dim=512, norm_first=True)
    loss_dict = dict(rollout_len=10, use_img_recon_loss=False)
    slot_recon_loss_w = 1.
    img_recon_loss_w = 1.
