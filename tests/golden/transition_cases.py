"""Seeded cases for the SAVi slot transition: predictor -> kernel_dist_layer -> sample of
reference StoSAVi.encode (savi.py:393-410).  Shared by make_golden_transition.py (unmodified reference on CPU) and
tests/test_transition.py (sm_100a kernel; it also tiles the clips of a case to larger batches -- clips are independent --
to reach every cluster size the launcher picks)."""
import numpy as np
import torch

from wrapper_cases import fill_seeded

_ENC = dict(enc_channels=(3, 64, 64, 64, 64), enc_ks=5, enc_out_channels=128, enc_norm='')
_DEC = dict(dec_channels=(128, 64, 64, 64, 64), dec_resolution=(8, 8), dec_ks=5, dec_norm='')


def _kw(K, pred, kld='none', kernel_mlp=True, mlp_size=256):
    return dict(resolution=(64, 64), clip_len=3,
                slot_dict=dict(num_slots=K, slot_size=128, slot_mlp_size=mlp_size, num_iterations=2, kernel_mlp=kernel_mlp),
                enc_dict=_ENC, dec_dict=_DEC, pred_dict=pred,
                loss_dict=dict(use_post_recon_loss=True, kld_method=kld))


def _tf(layers=2, heads=4, ffn=512, rnn=True, norm_first=True):
    return dict(pred_type='transformer', pred_rnn=rnn, pred_norm_first=norm_first, pred_num_layers=layers,
                pred_num_heads=heads, pred_ffn_dim=ffn, pred_sg_every=None)


# name -> (model kwargs, clips B, transitions after the first frame, weight seed)
CASES = {
    # savi_obj3d_params.py:38-72: Transformer (2 layers, 4 heads, ffn 4 D) -> LSTM(256) -> Linear, kernel MLP
    'tr_obj3d': (_kw(6, _tf()), 4, 3, 201),
    # stosavi_clevrer_params.py: residual MLP predictor, no RNN, single Linear head, stochastic kernels
    'tr_clevrer': (_kw(7, dict(pred_type='mlp', pred_rnn=False, pred_norm_first=True, pred_sg_every=None),
                       kld='var-0.01', kernel_mlp=False), 3, 2, 202),
    # post-LN Transformer, one layer, 8 heads, K = 8
    'tr_postln': (_kw(8, _tf(layers=1, heads=8, ffn=256, norm_first=False)), 5, 2, 203),
    # no RNN, K = 5, one clip
    'tr_plain': (_kw(5, _tf(rnn=False)), 1, 2, 204),
}


def build(cls, name):
    kw, B, steps, seed = CASES[name]
    torch.manual_seed(0)
    return fill_seeded(cls(**kw).eval(), seed)


def inputs(name):
    """prev_slots of every transition [steps, B, K, D] and the noise of every frame [steps + 1, B, K, D]."""
    kw, B, steps, seed = CASES[name]
    K, D = kw['slot_dict']['num_slots'], kw['slot_dict']['slot_size']
    rs = np.random.RandomState(seed + 1000)
    prev = rs.standard_normal((steps, B, K, D)).astype(np.float32)
    noise = rs.standard_normal((steps + 1, B, K, D)).astype(np.float32)
    return prev, noise
