"""Decoder epilogue: oracle vs the reference-generated goldens (CPU), and the sm_100a
kernels through the C ABI vs both (GPU).  Masks / colours are floating point (softmax in fp32: tolerance 2e-6
absolute, the spread between exp implementations); the segmentation is index arithmetic -> bit exact."""
import os
import sys

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, 'golden'))
import make_golden_decode as G  # noqa: E402
from oracle import slot_oracle as O  # noqa: E402

GOLD = np.load(os.path.join(HERE, 'golden', 'decode.npz'))


@pytest.mark.parametrize('name', list(G.CASES))
def test_oracle_matches_reference_decode_tail(name):
    x = G.make_input(*G.CASES[name])
    recon, masks = O.decode_combine(x)
    assert np.abs(masks - GOLD[name + '_masks']).max() < 2e-6
    assert np.abs(recon - GOLD[name + '_recon']).max() < 5e-6
    # postproc_mask is index arithmetic on given masks: bit exact against the unmodified reference
    seg = O.postproc_mask(GOLD[name + '_masks'][:, None])
    assert np.array_equal(seg, GOLD[name + '_seg'])


def test_oracle_postproc_ties_and_background_rule():
    m = np.zeros((1, 1, 3, 1, 1, 4), dtype=np.float32)
    m[0, 0, :, 0, 0] = [[0.4, 0.6, 0.2, 0.3], [0.4, 0.3, 0.2, 0.3], [0.2, 0.1, 0.6, 0.4]]
    # slot maxima 0.6, 0.4, 0.6 -> background = slot 1; pixels 0 and 3 have best score < 0.5 -> slot 1;
    # pixel 0's tie (0.4, 0.4) would have gone to slot 0 without the rule
    assert O.postproc_mask(m).reshape(-1).tolist() == [1, 0, 2, 1]


@pytest.mark.gpu
@pytest.mark.parametrize('name', list(G.CASES))
def test_kernel_matches_reference_and_oracle(name):
    from slotformer_b200 import engine
    x = torch.from_numpy(G.make_input(*G.CASES[name])).cuda()
    recon, masks, seg = engine.decode_combine(x, want_seg=True)
    assert masks.shape == GOLD[name + '_masks'].shape and recon.shape == GOLD[name + '_recon'].shape
    assert np.abs(masks.cpu().numpy() - GOLD[name + '_masks']).max() < 2e-6
    assert np.abs(recon.cpu().numpy() - GOLD[name + '_recon']).max() < 5e-6
    # index work: bit exact against the oracle run on the SAME masks, and (masks permitting) the reference
    assert np.array_equal(seg.cpu().numpy(), O.postproc_mask(masks.cpu().numpy()[:, None])[:, 0])
    # stand-alone entry point on the reference's masks: bit exact against the reference's own result
    from slotformer_b200.video_prediction.vp_utils import postproc_mask
    ref_masks = torch.from_numpy(GOLD[name + '_masks']).cuda().unsqueeze(1)
    assert np.array_equal(postproc_mask(ref_masks).cpu().numpy(), GOLD[name + '_seg'])


@pytest.mark.gpu
def test_kernel_full_size_properties():
    """OBJ3D decode size (64 frames x 6 slots x 128x128): masks sum to 1, colours stay inside the slot hull,
    slot permutation permutes masks and relabels the segmentation, repeat runs are bit identical."""
    from slotformer_b200 import engine
    gen = torch.Generator(device='cuda').manual_seed(3)
    x = torch.randn((64, 6, 4, 128, 128), device='cuda', generator=gen)
    x[:, :, 3] *= 5
    recon, masks, seg = engine.decode_combine(x, want_seg=True)
    assert torch.allclose(masks.sum(1), torch.ones_like(masks[:, 0]), atol=1e-6)
    assert (recon <= x[:, :, :3].amax(1) + 1e-5).all() and (recon >= x[:, :, :3].amin(1) - 1e-5).all()
    r2, m2, s2 = engine.decode_combine(x, want_seg=True)
    assert torch.equal(recon, r2) and torch.equal(masks, m2) and torch.equal(seg, s2)
    perm = torch.tensor([3, 0, 5, 1, 4, 2], device='cuda')
    rp, mp, sp = engine.decode_combine(x[:, perm].contiguous(), want_seg=True)
    assert torch.allclose(mp, masks[:, perm], atol=3e-7)      # the softmax denominator is summed in slot order
    assert torch.allclose(rp, recon, atol=1e-5)
    inv = torch.empty_like(perm)
    inv[perm] = torch.arange(6, device='cuda')
    top2 = masks.squeeze(2).topk(2, dim=1).values
    # pixels decided by their own arg-max (best score above the threshold, no near tie); the background slot of a
    # frame is a first-minimum over slot maxima that saturate at 1.0 here, so it legitimately depends on slot order
    clear = ((top2[:, 0] - top2[:, 1]) > 1e-5) & (top2[:, 0] > 0.5 + 1e-5)
    assert clear.float().mean() > 0.5
    assert torch.equal(inv[seg][clear], sp[clear])


@pytest.mark.gpu
def test_decode_rejects_bad_input():
    from slotformer_b200 import engine
    with pytest.raises(engine.SfbError):
        engine.decode_combine(torch.zeros((1, 3, 4, 4, 4)))            # CPU tensor
    with pytest.raises(engine.SfbError):
        engine.decode_combine(torch.zeros((1, 3, 3, 4, 4), device='cuda'))   # 3 planes
    r, m = engine.decode_combine(torch.zeros((0, 3, 4, 4, 4), device='cuda'))
    assert r.shape == (0, 3, 4, 4) and m.shape == (0, 3, 1, 4, 4)


@pytest.mark.gpu
def test_postproc_mask_large_batches_and_negative_scores():
    """More (frame, slot) planes than one grid dimension holds (64 x 160 x 7 = 71680 > 65535) and scores of either
    sign (postproc_mask takes arbitrary scores, vp_utils.py:20-41): bit-exact against the oracle."""
    from slotformer_b200 import engine
    gen = torch.Generator(device='cuda').manual_seed(9)
    m = torch.randn((64, 160, 7, 1, 8, 8), device='cuda', generator=gen)          # mostly below the 0.5 threshold
    seg = engine.postproc_mask(m)
    assert seg.shape == (64, 160, 8, 8)
    ref = O.postproc_mask(m.cpu().numpy())
    assert np.array_equal(seg.cpu().numpy(), ref)
    neg = -torch.rand((3, 2, 4, 1, 4, 4), device='cuda', generator=gen) - 0.1      # every score negative
    assert np.array_equal(engine.postproc_mask(neg).cpu().numpy(), O.postproc_mask(neg.cpu().numpy()))
