"""SAVi slot transition: predictor -> kernel_dist_layer -> sample as one sm_100a kernel
(csrc/transition.cu) against the UNMODIFIED reference StoSAVi's chain (tests/golden/transition.npz, produced by
tests/golden/make_golden_transition.py: savi.py:393-410, predictor.py:20-113 evaluated in fp64)."""
import ctypes
import os

import numpy as np
import pytest
import torch

import transition_cases as TC
from helpers import GOLD, rel_max

# fp32 FFMA chains in another summation order than the reference's: the reference's own fp32 evaluation sits
# 2e-7 ... 1.2e-6 (max-norm, relative) from its fp64 evaluation (the *_f32err entries of the golden file)
TOL = 5e-6


def _gold():
    return np.load(os.path.join(GOLD, 'transition.npz'))


# ---------------------------------------------------------------------------------------------- CPU
def test_goldens_cover_every_case_and_state():
    g = _gold()
    for name, (kw, B, steps, _) in TC.CASES.items():
        K, D = kw['slot_dict']['num_slots'], kw['slot_dict']['slot_size']
        assert g[f'{name}.dist'].shape == (steps + 1, B, K, 2 * D)
        assert g[f'{name}.init'].shape == (steps + 1, B, K, D)
        assert float(g[f'{name}.dist_f32err']) < 2e-6
        if kw['pred_dict']['pred_rnn']:
            assert g[f'{name}.h'].shape == (1, B * K, kw['slot_dict']['slot_mlp_size'])
        if kw['loss_dict']['kld_method'] == 'none':      # deterministic SAVi: the kernels are the means
            assert np.array_equal(g[f'{name}.init'], g[f'{name}.dist'][..., :D])


def test_transition_spec_and_workspace_on_cpu():
    """Host logic without a GPU: the structure description of every case, the C ABI's layout size, and the
    rejection of structures outside the kernel's envelope."""
    from slotformer_b200 import engine
    from slotformer_b200.base_slots.models import StoSAVi
    lib = engine.load()
    for name, (kw, B, steps, _) in TC.CASES.items():
        m = TC.build(StoSAVi, name)
        spec = m._transition_spec()
        assert spec is not None, name
        cw, _keep = None, None
        # CPU tensors are refused by the launcher, the layout query needs no device
        with pytest.raises(engine.SfbError):
            engine.TransitionEngine._pack(spec)
        cw = engine._TRWeights()
        cw.pred_type, cw.num_layers, cw.num_heads = spec['pred_type'], spec['num_layers'], spec['num_heads']
        cw.ffn_dim, cw.norm_first, cw.mlp_hidden = spec['ffn_dim'], int(spec['norm_first']), spec['mlp_hidden']
        cw.rnn_hidden, cw.kernel_mlp = spec['rnn_hidden'], int(spec['kernel_mlp'])
        nbytes = int(lib.sfb_transition_workspace_bytes(ctypes.byref(cw), 128))
        nparam = sum(int(np.prod(t.shape)) for t in spec['weights'].values())
        if spec['rnn_hidden']:
            nparam -= 4 * spec['rnn_hidden']             # b_ih + b_hh are stored as one vector
        assert nbytes == 4 * nparam, (name, nbytes, nparam)
        cw.ffn_dim = 2048                                # wider than the kernel's activation buffers
        if spec['pred_type'] == engine.SFB_TR_TRANSFORMER:
            assert int(lib.sfb_transition_workspace_bytes(ctypes.byref(cw), 128)) == 0
    # a GRU cell is not covered: the module keeps the stock path
    m = TC.build(StoSAVi, 'tr_obj3d')
    m.predictor.rnn = torch.nn.GRU(128, 256)
    assert m._transition_spec() is None


# ---------------------------------------------------------------------------------------------- GPU
def _run_case(name, tile):
    from slotformer_b200 import engine
    from slotformer_b200.base_slots.models import StoSAVi
    kw, B0, steps, _ = TC.CASES[name]
    g = _gold()
    dev = 'cuda:0'
    m = TC.build(StoSAVi, name).to(dev)
    spec = m._transition_spec()
    assert spec is not None and engine.TransitionEngine.supported(spec, m.slot_size)
    prev, noise = TC.inputs(name)
    B = B0 * tile
    rep = lambda a, axis: np.concatenate([a] * tile, axis=axis)      # clips are independent: tile the batch
    stochastic = m.kld_method != 'none'
    eng = engine.TransitionEngine()
    state = None
    n0 = engine.launch_count()
    errs = []
    for t in range(steps + 1):
        nz = torch.from_numpy(rep(noise[t], 0)).to(dev) if stochastic else None
        if t == 0:
            dist, init, new = eng.forward(spec, m.init_latents.detach(), False, B, None, nz)
            assert new is None
        else:
            p = torch.from_numpy(rep(prev[t - 1], 0)).to(dev)
            dist, init, new = eng.forward(spec, p, True, B, state, nz)
            state = new
        errs.append(rel_max(dist.cpu().numpy(), rep(g[f'{name}.dist'][t], 0)))
        errs.append(rel_max(init.cpu().numpy(), rep(g[f'{name}.init'][t], 0)))
    assert engine.launch_count() - n0 == steps + 1          # ONE kernel per frame
    if spec['rnn_hidden']:
        K = kw['slot_dict']['num_slots']
        h = g[f'{name}.h'][0].reshape(B0, K, -1)
        c = g[f'{name}.c'][0].reshape(B0, K, -1)
        errs.append(rel_max(state[0].cpu().numpy().reshape(B, K, -1), rep(h, 0)))
        errs.append(rel_max(state[1].cpu().numpy().reshape(B, K, -1), rep(c, 0)))
    else:
        assert state is None
    return max(errs)


@pytest.mark.gpu
@pytest.mark.parametrize('name', list(TC.CASES))
def test_transition_matches_reference(name):
    assert _run_case(name, 1) < TOL


@pytest.mark.gpu
@pytest.mark.parametrize('name,tile', [('tr_obj3d', 5), ('tr_obj3d', 10), ('tr_obj3d', 40), ('tr_postln', 6),
                                       ('tr_clevrer', 30), ('tr_plain', 19)])
def test_transition_every_cluster_size(name, tile):
    """B * 8 <= 148 SMs runs 8-CTA clusters per clip, then 4, 2 and 1 CTA per clip as the batch grows."""
    assert _run_case(name, tile) < TOL


@pytest.mark.gpu
def test_transition_is_deterministic_and_batch_independent():
    from slotformer_b200 import engine
    from slotformer_b200.base_slots.models import StoSAVi
    dev = 'cuda:0'
    m = TC.build(StoSAVi, 'tr_obj3d').to(dev)
    spec = m._transition_spec()
    eng = engine.TransitionEngine()
    torch.manual_seed(3)
    prev = torch.randn(24, 6, 128, device=dev)
    h = torch.randn(24 * 6, 256, device=dev)
    c = torch.randn(24 * 6, 256, device=dev)
    a = eng.forward(spec, prev, True, 24, (h, c))
    b = eng.forward(spec, prev, True, 24, (h, c))
    for x, y in zip((a[0], a[1], *a[2]), (b[0], b[1], *b[2])):
        assert torch.equal(x, y)
    # a sub-batch (other cluster size, other CTA -> feature mapping) gives the same values: per output element the
    # k-slice partial sums are combined in a fixed order that does not depend on the split
    s = eng.forward(spec, prev[:3].contiguous(), True, 3, (h[:18].contiguous(), c[:18].contiguous()))
    assert rel_max(s[0].cpu().numpy(), a[0][:3].cpu().numpy()) < 1e-6
    assert rel_max(s[2][1].cpu().numpy(), a[2][1][:18].cpu().numpy()) < 1e-6


@pytest.mark.gpu
def test_savi_encode_fused_transition_matches_stock_modules():
    """StoSAVi.encode with the fused transition == the same model running predictor / kernel_dist_layer as stock
    PyTorch modules (both feed the same Slot Attention kernels), eager and as a CUDA-graph replay."""
    import wrapper_cases as W
    from slotformer_b200.base_slots.models import StoSAVi
    dev = 'cuda:0'
    m = W.build_savi(StoSAVi).to(dev)
    img = W.savi_input().to(dev)
    outs = {}
    with torch.no_grad():
        for fused in (False, True):
            for graph in (False, True):
                m.fuse_transition, m.use_cuda_graph = fused, graph
                m.__dict__.pop('_loop_graphs', None)
                m._reset_rnn()
                d1, s1, _ = m.encode(img)
                d2, s2, _ = m.encode(img, prev_slots=s1[:, -1])       # carries slots and the LSTM state on
                outs[(fused, graph)] = [x.cpu().numpy() for x in (d1, s1, d2, s2)]
    ref = outs[(False, False)]
    for key, val in outs.items():
        for a, b in zip(val, ref):
            assert rel_max(a, b) < 5e-4, key         # through two Slot Attention iterations per frame (fp16 operands)
    for a, b in zip(outs[(True, True)], outs[(True, False)]):
        assert np.array_equal(a, b)                                     # the replay runs the same kernels
