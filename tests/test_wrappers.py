"""Caller modules (SURVEY.md section 8 a9 / 8 b): StoSAVi and SlotFormer keep the reference's
state_dict keys and reproduce the reference's outputs (tests/golden/wrappers.npz, produced by
the UNMODIFIED reference loaded with our seeded state_dict)."""
import glob
import importlib.util
import os
import tempfile

import numpy as np
import pytest
import torch

import wrapper_cases as W
from helpers import GOLD, rel_max

REF = '/root/reference/slotformer'


def _gold():
    return np.load(os.path.join(GOLD, 'wrappers.npz'))


def test_state_dict_keys_match_reference():
    from slotformer_b200.base_slots.models import StoSAVi
    from slotformer_b200.video_prediction.models import SlotFormer
    g = _gold()
    savi = W.build_savi(StoSAVi)
    assert list(savi.state_dict().keys()) == g['savi_keys'].tolist()
    with tempfile.TemporaryDirectory() as td:
        ckpt = os.path.join(td, 'savi.pth')
        torch.save({'state_dict': savi.state_dict()}, ckpt)
        sf = W.build_slotformer(SlotFormer, ckpt)
    assert list(sf.state_dict().keys()) == g['sf_keys'].tolist()
    # the decoder was sliced out of the SAVi checkpoint by key prefix and is frozen
    assert all(not p.requires_grad for p in sf.decoder.parameters())
    assert torch.equal(sf.decoder[0][0].weight, savi.decoder[0][0].weight)


@pytest.mark.skipif(not os.path.isdir(REF), reason='reference config files only exist in the build container')
def test_build_model_with_every_reference_config():
    """params.py files of the reference load unchanged through the nerv shim and build."""
    from slotformer_b200.compat import install_nerv_shim
    from slotformer_b200 import base_slots, video_prediction
    install_nerv_shim()

    def load(path):
        spec = importlib.util.spec_from_file_location('cfg_' + os.path.basename(path)[:-3].replace('-', '_'), path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        return mod.SlotFormerParams()

    built = 0
    with tempfile.TemporaryDirectory() as td:
        ckpts = {}
        for path in sorted(glob.glob(f'{REF}/base_slots/configs/*savi*_params*.py')):
            params = load(path)
            assert params.get('model') == 'StoSAVi' and params.get('no_such_key', 5) == 5
            model = base_slots.build_model(params)
            key = (tuple(params.resolution), params.slot_dict['slot_size'], tuple(params.dec_dict['dec_channels']),
                   tuple(params.dec_dict['dec_resolution']))
            ckpts[key] = os.path.join(td, f'savi{len(ckpts)}.pth')
            torch.save({'state_dict': model.state_dict()}, ckpts[key])
            built += 1
        for path in sorted(glob.glob(f'{REF}/video_prediction/configs/slotformer_*_params*.py')):
            params = load(path)
            if params.model == 'STEVESlotFormer':
                model = video_prediction.build_model(params)          # rollout only: no dVAE / decoder checkpoint needed
                assert model.rollouter.history_len == params.rollout_dict['history_len'] and model.decoder is None
                built += 1
                continue
            key = (tuple(params.resolution), params.slot_dict['slot_size'], tuple(params.dec_dict['dec_channels']),
                   tuple(params.dec_dict['dec_resolution']))
            assert key in ckpts, f'no SAVi decoder for {path}'
            params.dec_dict['dec_ckp_path'] = ckpts[key]
            model = video_prediction.build_model(params)
            assert model.rollouter.history_len == params.rollout_dict['history_len']
            built += 1
        for path in sorted(glob.glob(f'{REF}/base_slots/configs/steve_*_params*.py')):
            params = load(path)
            assert params.model == 'STEVE'
            model = base_slots.build_model(params)                    # slot-extraction half, no dVAE checkpoint needed
            assert type(model.slot_attention).__name__ == 'SlotAttentionWMask'
            built += 1
    assert built >= 8


@pytest.mark.gpu
def test_stosavi_matches_reference_on_gpu():
    from slotformer_b200.base_slots.models import StoSAVi
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    g = _gold()
    m = W.build_savi(StoSAVi).cuda()
    img = W.savi_input().cuda()
    with torch.no_grad():
        m.testing = True
        out = m({'img': img})
        assert set(out) == {'post_slots', 'kernel_dist', 'img'}
        assert rel_max(out['post_slots'].cpu().numpy(), g['savi_post_slots']) < 2e-3
        m.testing = False
        full = m({'img': img})
    assert rel_max(full['post_recon_combined'].cpu().numpy(), g['savi_recon']) < 2e-3
    assert np.abs(full['post_masks'].cpu().numpy() - g['savi_masks']).max() < 2e-3
    loss = m.calc_train_loss({'img': img}, full)
    assert set(loss) == {'kld_loss', 'post_recon_loss'}


@pytest.mark.gpu
def test_slotformer_matches_reference_on_gpu():
    from slotformer_b200.base_slots.models import StoSAVi
    from slotformer_b200.video_prediction.models import SlotFormer
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    g = _gold()
    with tempfile.TemporaryDirectory() as td:
        ckpt = os.path.join(td, 'savi.pth')
        torch.save({'state_dict': W.build_savi(StoSAVi).state_dict()}, ckpt)
        m = W.build_slotformer(SlotFormer, ckpt).cuda()
    x = W.slotformer_input().cuda()
    with torch.no_grad():
        fwd = m({'slots': x})
        assert rel_max(fwd['pred_slots'].cpu().numpy(), g['sf_pred']) < 4e-3
        assert torch.equal(fwd['gt_slots'], x[:, 6:])
        dec = m.rollout(x[:, :6], 3, decode=True, with_gt=False)
        assert rel_max(dec['recon_combined'].cpu().numpy(), g['sf_recon']) < 4e-3
        loss = m.calc_train_loss({'slots': x}, fwd)
    assert abs(loss['slot_recon_loss'].item() - float(g['sf_loss'])) < 1e-2 * float(g['sf_loss'])
    m.rollout_len = 2          # mutable at run time, as the offline rollout scripts do
    with torch.no_grad():
        assert m({'slots': x[:, :8]})['pred_slots'].shape == (2, 2, 5, 128)


@pytest.mark.gpu
def test_training_step_uses_autograd_path():
    """Gradients flow through both operators (differentiable restatement, GPU eager)."""
    from slotformer_b200.base_slots.models import SlotAttention
    from slotformer_b200.video_prediction.models import SlotRollouter
    sa = SlotAttention(128, 2, 4, 128, 256).cuda().train()
    out = sa(torch.randn(2, 256, 128, device='cuda'), torch.randn(2, 4, 128, device='cuda'))
    out.square().mean().backward()
    assert sa.project_k.weight.grad is not None and torch.isfinite(sa.project_k.weight.grad).all()
    ro = SlotRollouter(3, 128, 2, num_layers=2).cuda().eval()
    x = torch.randn(2, 2, 3, 128, device='cuda', requires_grad=True)
    ro(x, 2).square().mean().backward()
    assert x.grad is not None and ro.in_proj.weight.grad is not None


@pytest.mark.gpu
def test_savi_frame_loop_cuda_graph_matches_eager():
    """the per-frame chain (predictor -> distribution head -> Slot Attention) replayed as one
    CUDA graph gives bit-identical slots to the eager loop, for the first clip (no previous slots) and for a
    continuation (previous slots given), and is re-captured when a weight changes."""
    from slotformer_b200.base_slots.models import StoSAVi
    m = W.build_savi(StoSAVi).cuda()
    img = torch.cat([W.savi_input()] * 2, dim=0).cuda()          # [4, 3, 3, 64, 64]
    def both(x, graph):
        """first clip (fresh recurrent state), then a continuation that carries slots AND the LSTM state over"""
        m.use_cuda_graph = graph
        m.predictor.reset()
        d0, s0, _ = m.encode(x)
        d1, s1, _ = m.encode(x, prev_slots=s0[:, -1])
        return d0, s0, d1, s1

    with torch.no_grad():
        ref = both(img, False)
        for _ in range(2):                                       # capture, then replay
            got = both(img, True)
            assert all(torch.equal(a, b) for a, b in zip(ref, got))
        assert m.use_cuda_graph and len(m._loop_graphs) >= 2     # really went through the graphs
        img2 = img.flip(0).contiguous()                          # new inputs through the captured graphs
        assert all(torch.equal(a, b) for a, b in zip(both(img2, False), both(img2, True)))
        m.slot_attention.project_q[1].weight.mul_(1.01)          # a changed weight invalidates the capture
        new = both(img2, True)
        assert all(torch.equal(a, b) for a, b in zip(both(img2, False), new))
        assert not torch.equal(new[1], got[1])


def test_encode_features_split_equals_encode_on_cpu():
    """``encode_features`` + ``encode(feats=...)`` (the two stages a caller may put on different streams) == ``encode``;
    on the CPU both run the stock layers and the differentiable Slot Attention restatement."""
    from slotformer_b200.base_slots.models import StoSAVi
    m = W.build_savi(StoSAVi)
    img = W.savi_input()[:1, :2]
    with torch.enable_grad():                                   # CPU tensors only run through the autograd restatement
        m.predictor.reset()
        d0, s0, f0 = m.encode(img)
        m.predictor.reset()
        feats = m.encode_features(img)
        d1, s1, f1 = m.encode(None, feats=feats)
    assert torch.equal(f0, f1) and torch.equal(d0, d1) and torch.equal(s0, s1)
    assert feats.shape == (1, 2, 4096, 128)


def test_bind_to_gpu_numa_is_harmless_without_a_gpu():
    import os
    from slotformer_b200.parallel import bind_to_gpu_numa
    before = os.sched_getaffinity(0)
    cpus = bind_to_gpu_numa(0)
    assert cpus is None or set(cpus) <= before
    if cpus is None:
        assert os.sched_getaffinity(0) == before
    else:
        os.sched_setaffinity(0, before)


@pytest.mark.gpu
def test_savi_graphs_survive_workspace_growth():
    """a captured frame loop holds the engines' workspace pointers.  B = 2, then B = 8 (the
    workspaces are re-allocated), then B = 2 again replays the OLD graph: its workspaces must still be alive and its
    results must equal the eager loop."""
    from slotformer_b200.base_slots.models import StoSAVi
    m = W.build_savi(StoSAVi).cuda()
    small = W.savi_input().cuda()                                   # [2, 3, 3, 64, 64]
    big = torch.cat([small, small.flip(0), small * 0.5, small.flip(1)], dim=0).contiguous()

    def run(x, graph):
        m.use_cuda_graph = graph
        m.predictor.reset()
        return m.encode(x)[1]

    with torch.no_grad():
        ref_small, ref_big = run(small, False), run(big, False)
        a = run(small, True)
        b = run(big, True)
        junk = [torch.full((1 << 22,), float('nan'), device='cuda') for _ in range(8)]   # reuse of any freed block shows
        c = run(small, True)
        del junk
        assert torch.equal(a, ref_small) and torch.equal(b, ref_big) and torch.equal(c, ref_small)


def _steve_gold():
    return np.load(os.path.join(os.path.dirname(__file__), 'golden', 'steve.npz'))


def test_steve_callers_keep_reference_state_dict_keys():
    """STEVE (slot extraction) and STEVESlotFormer (rollout only) build everything the reference builds apart from
    the dVAE / token decoder, under the reference's state_dict keys; a reference checkpoint loads strictly once its
    dvae.* / trans_decoder.* / decoder.* entries are dropped (which load_state_dict does)."""
    from slotformer_b200.base_slots.models import STEVE
    from slotformer_b200.video_prediction.models import STEVESlotFormer
    g = _steve_gold()
    m = W.build_steve(STEVE)
    assert sorted(m.state_dict().keys()) == sorted(g['steve_keys'].tolist())
    sd = dict(m.state_dict())
    sd['dvae.encoder.0.weight'] = torch.zeros(1)
    sd['trans_decoder.head.weight'] = torch.zeros(1)
    m.load_state_dict(sd, strict=True)
    sf = W.build_steve_slotformer(STEVESlotFormer)
    assert sorted(sf.state_dict().keys()) == sorted(g['ssf_keys'].tolist())
    with pytest.raises(NotImplementedError):
        sf.decode(torch.zeros(1, 6, 192))


@pytest.mark.gpu
def test_steve_extraction_matches_reference_on_gpu():
    """BASELINE config 4 caller: STEVE.encode through SlotAttentionWMask (C = D = 192, masks up-sampled to 128x128)."""
    from slotformer_b200.base_slots.models import STEVE
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    g = _steve_gold()
    m = W.build_steve(STEVE).cuda()
    img = W.steve_input().cuda()
    with torch.no_grad():
        out = m({'img': img})
    assert set(out) == {'slots', 'masks'}
    assert out['masks'].shape == (2, 3, 6, 128, 128)
    assert rel_max(out['slots'].cpu().numpy(), g['steve_slots']) < 2e-3
    assert np.abs(out['masks'].cpu().numpy()[..., 1::4, 2::4] - g['steve_masks_sub']).max() < 2e-3


@pytest.mark.gpu
def test_steve_slotformer_rollout_matches_reference_on_gpu():
    from slotformer_b200.video_prediction.models import STEVESlotFormer
    g = _steve_gold()
    m = W.build_steve_slotformer(STEVESlotFormer).cuda()
    x = W.steve_slotformer_input().cuda()
    with torch.no_grad():
        fwd = m({'slots': x})
        loss = m.calc_train_loss({'slots': x}, fwd)
    assert rel_max(fwd['pred_slots'].cpu().numpy(), g['ssf_pred']) < 4e-3
    assert torch.equal(fwd['gt_slots'], x[:, 6:])
    assert abs(loss['slot_recon_loss'].item() - float(g['ssf_loss'])) < 1e-2 * float(g['ssf_loss'])
