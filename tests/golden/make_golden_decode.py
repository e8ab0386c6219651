#!/usr/bin/env python
"""Goldens for the decoder epilogue: the reference's own expressions for the tail of
StoSAVi.decode (savi.py:519-523, executed verbatim with torch) and the UNMODIFIED reference postproc_mask
(video_prediction/vp_utils.py:20-41).  Build-container only (needs /root/reference)."""
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, HERE)
import make_golden  # noqa: E402  (reference import machinery)

CASES = {  # name: (B, K, H, W, logit scale, seed)
    'dec_tiny': (2, 5, 8, 8, 3.0, 5),
    'dec_obj3d': (3, 6, 64, 64, 6.0, 6),        # OBJ3D decode size / 2, peaky masks
    'dec_flat': (2, 7, 16, 16, 0.05, 7),        # nearly uniform masks: every pixel below FG_THRE
}


def make_input(B, K, H, W, scale, seed):
    rs = np.random.RandomState(seed)
    x = rs.standard_normal((B, K, 4, H, W)).astype(np.float32)
    x[:, :, 3] *= scale
    return x


def main():
    make_golden.import_reference()
    from slotformer.video_prediction.vp_utils import postproc_mask
    out = {}
    for name, (B, K, H, W, scale, seed) in CASES.items():
        x = torch.from_numpy(make_input(B, K, H, W, scale, seed))
        recons = x[:, :, :3, :, :]                      # savi.py:519
        masks = x[:, :, -1:, :, :]                      # savi.py:520
        masks = F.softmax(masks, dim=1)                 # savi.py:521
        recon_combined = torch.sum(recons * masks, dim=1)   # savi.py:522
        seg = postproc_mask(masks.unsqueeze(1))         # [B, 1, K, 1, H, W] -> [B, 1, H, W]
        out[name + '_masks'] = masks.numpy()
        out[name + '_recon'] = recon_combined.numpy()
        out[name + '_seg'] = seg.numpy().astype(np.int64)
    path = os.path.join(HERE, 'decode.npz')
    np.savez_compressed(path, **out)
    print('wrote', path, os.path.getsize(path) // 1024, 'KiB')


if __name__ == '__main__':
    main()
