#!/usr/bin/env python
"""Goldens for the caller modules: the UNMODIFIED reference StoSAVi /
SlotFormer, loaded with the state_dict of OUR seeded models (strict -> proves key compatibility),
run on CPU.  Build-container only (needs /root/reference)."""
import os
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, HERE)
import make_golden  # noqa: E402  (reference import machinery)
import wrapper_cases as W  # noqa: E402


def main():
    make_golden.import_reference()
    from slotformer.base_slots.models import StoSAVi as RefSAVi
    from slotformer.video_prediction.models import SlotFormer as RefSlotFormer
    from slotformer_b200.base_slots.models import StoSAVi
    from slotformer_b200.video_prediction.models import SlotFormer

    ours = W.build_savi(StoSAVi)
    ref = RefSAVi(**W.SAVI_KW).eval()
    ref.load_state_dict(ours.state_dict(), strict=True)
    img = W.savi_input()
    with torch.no_grad():
        ref.testing = True
        slots = ref({'img': img})['post_slots']
        ref.testing = False
        full = ref({'img': img})
    out = {'savi_keys': np.array(list(ref.state_dict().keys())), 'savi_post_slots': slots.numpy(),
           'savi_recon': full['post_recon_combined'].numpy(), 'savi_masks': full['post_masks'].numpy()}

    with tempfile.TemporaryDirectory() as td:
        ckpt = os.path.join(td, 'savi.pth')
        torch.save({'state_dict': ours.state_dict()}, ckpt)
        ours_sf = W.build_slotformer(SlotFormer, ckpt)
        kw = dict(W.SLOTFORMER_KW)
        kw['dec_dict'] = dict(kw['dec_dict'], dec_ckp_path=ckpt)
        ref_sf = RefSlotFormer(**kw).eval()
        ref_sf.load_state_dict(ours_sf.state_dict(), strict=True)
    x = W.slotformer_input()
    with torch.no_grad():
        fwd = ref_sf({'slots': x})
        dec = ref_sf.rollout(x[:, :6], 3, decode=True, with_gt=False)
        losses = ref_sf.calc_train_loss({'slots': x}, fwd)
    out.update(sf_keys=np.array(list(ref_sf.state_dict().keys())), sf_pred=fwd['pred_slots'].numpy(),
               sf_recon=dec['recon_combined'].numpy(), sf_loss=np.float64(losses['slot_recon_loss'].item()))
    path = os.path.join(HERE, 'wrappers.npz')
    np.savez_compressed(path, **out)
    print('wrote', path, os.path.getsize(path) // 1024, 'KiB')


if __name__ == '__main__':
    main()
