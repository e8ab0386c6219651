#!/usr/bin/env python
"""Goldens for the STEVE callers (BASELINE config 4): the UNMODIFIED reference STEVE
(slot extraction, `testing = True`) and STEVESlotFormer.rollout on CPU.  The reference needs a dVAE and a
decoder checkpoint to even construct (steve.py:169-172, steve_slotformer.py:76-80): both are minted here from
seeded reference instances.  Weights of everything our models build are drawn per state_dict key from a numpy
stream (wrapper_cases.fill_seeded), so the test re-creates them without storing them; the key lists are stored
to prove key compatibility.  Build-container only."""
import os
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, HERE)
import make_golden  # noqa: E402
import wrapper_cases as W  # noqa: E402


def main():
    make_golden.import_reference()
    from slotformer.base_slots.models import STEVE as RefSTEVE, dVAE as RefDVAE
    from slotformer.video_prediction.models import STEVESlotFormer as RefSSF
    out = {}
    with tempfile.TemporaryDirectory() as td:
        torch.manual_seed(20)
        dvae = RefDVAE(vocab_size=W.STEVE_KW['dvae_dict']['vocab_size'], img_channels=3)
        dvae_ckpt = os.path.join(td, 'dvae.pth')
        torch.save({'state_dict': dvae.state_dict()}, dvae_ckpt)
        ref = W.build_steve(RefSTEVE, dvae_dict=dict(W.STEVE_KW['dvae_dict'], dvae_ckp_path=dvae_ckpt))
        ref.testing = True
        img = W.steve_input()
        with torch.no_grad():
            r = ref({'img': img})
        sd = ref.state_dict()
        out['steve_keys'] = np.array([k for k in sd if not k.startswith(('dvae.', 'trans_decoder.'))])
        out['steve_slots'] = r['slots'].numpy()
        out['steve_masks_sub'] = r['masks'].numpy()[..., 1::4, 2::4]      # [2,3,6,32,32] of the up-sampled 128x128 masks
        steve_ckpt = os.path.join(td, 'steve.pth')
        torch.save({'state_dict': sd}, steve_ckpt)
        ref_sf = W.build_steve_slotformer(
            RefSSF, dvae_dict=dict(W.STEVE_SLOTFORMER_KW['dvae_dict'], dvae_ckp_path=dvae_ckpt),
            dec_dict=dict(W.STEVE_SLOTFORMER_KW['dec_dict'], dec_ckp_path=steve_ckpt))
        x = W.steve_slotformer_input()
        with torch.no_grad():
            fwd = ref_sf({'slots': x})
            loss = ref_sf.calc_train_loss({'slots': x}, fwd)
        sd2 = ref_sf.state_dict()
        out['ssf_keys'] = np.array([k for k in sd2 if not k.startswith(('dvae.', 'decoder.'))])
        out['ssf_pred'] = fwd['pred_slots'].numpy()
        out['ssf_loss'] = np.float64(loss['slot_recon_loss'].item())
    path = os.path.join(HERE, 'steve.npz')
    np.savez_compressed(path, **out)
    print('wrote', path, os.path.getsize(path) // 1024, 'KiB')


if __name__ == '__main__':
    main()
