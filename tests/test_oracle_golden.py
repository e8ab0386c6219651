"""Pins oracle/slot_oracle.py against outputs of the UNMODIFIED reference
modules (tests/golden/*.npz, produced by tests/golden/make_golden.py from
/root/reference: savi.py:56-102, steve.py:43-73, slotformer.py:85-126,
single_step_slotformer.py:49-90)."""
import os

import numpy as np
import pytest

import cases
from oracle import slot_oracle as O

GOLD = os.path.join(os.path.dirname(__file__), 'golden')


def _rel(a, b):
    return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-30))


@pytest.mark.parametrize('name', list(cases.SA_CASES))
def test_slot_attention_matches_reference(name):
    c, w, feats, slots = cases.sa_case(name)
    g = np.load(os.path.join(GOLD, f'{name}.npz'))
    out = O.slot_attention(feats, slots, w, c['iters'], return_mask=c['mask'])
    if c['mask']:
        out, mask = out
        assert mask.shape == (c['B'], c['K'], c['N'])
        assert _rel(mask, g['mask_f64']) < 1e-9
        assert _rel(mask, g['mask_f32']) < 1e-4
    assert out.shape == (c['B'], c['K'], c['D'])
    assert _rel(out, g['slots_f64']) < 1e-9          # fp64 oracle == fp64 reference
    assert _rel(out, g['slots_f32']) < 1e-4          # fp32 reference rounding


@pytest.mark.parametrize('name', list(cases.RO_CASES))
def test_rollout_matches_reference(name):
    c, w, hist = cases.ro_case(name)
    g = np.load(os.path.join(GOLD, f'{name}.npz'))
    w = dict(w)
    w['enc_t_pe'] = g['enc_t_pe']
    out = O.rollout(hist, w, c['pred_len'], c['heads'], c['layers'],
                    mode=c['mode'], cond_len=c['cond_len'])
    assert out.shape == (c['B'], c['pred_len'], c['K'], c['Ds'])
    assert _rel(out, g['pred_f64']) < 1e-8
    assert _rel(out, g['pred_f32']) < 2e-3           # fp32 drift over up to 64 AR steps


@pytest.mark.parametrize('L,d', [(6, 128), (15, 256), (1, 128)])
def test_sin_pos_enc_matches_reference_table(L, d):
    # enc_t_pe stored by the reference (slotformer.py:10-16) for the same shape
    for name, c in cases.RO_CASES.items():
        pe_len = c['cond_len'] if c['mode'] == 'grow' else c['T_h']
        if pe_len == L and c['d'] == d:
            g = np.load(os.path.join(GOLD, f'{name}.npz'))
            np.testing.assert_allclose(O.sin_pos_enc(L, d, np.float32)[0],
                                       g['enc_t_pe'][0], rtol=0, atol=2e-6)
            return
    pe = O.sin_pos_enc(L, d)
    assert pe.shape == (1, L, d)
    assert np.allclose(pe[0, -1, :d // 2], 0.0) and np.allclose(pe[0, -1, d // 2:], 1.0)


def test_state_dict_keys_match_reference():
    g = np.load(os.path.join(GOLD, 'sa_tiny.npz'))
    _, w, _, _ = cases.sa_case('sa_tiny')
    assert set(g['keys'].tolist()) == set(w)
    g = np.load(os.path.join(GOLD, 'ro_tiny.npz'))
    _, w, _ = cases.ro_case('ro_tiny')
    assert set(g['keys'].tolist()) == set(w) | {'enc_t_pe'}


@pytest.mark.parametrize('name', ['tr_obj3d', 'tr_clevrer', 'tr_postln', 'tr_plain'])
def test_transition_oracle_matches_reference(name):
    """The numpy restatement of the SAVi slot transition (predictor -> kernel_dist_layer -> sample) against the
    unmodified reference StoSAVi's chain (tests/golden/transition.npz: fp64 evaluation stored as fp32)."""
    import os
    import transition_cases as TC
    from helpers import GOLD
    from slotformer_b200.base_slots.models import StoSAVi
    kw, B, steps, _ = TC.CASES[name]
    g = np.load(os.path.join(GOLD, 'transition.npz'))
    m = TC.build(StoSAVi, name)                      # the seeded weights the goldens were made with
    w = {k: v.numpy() for k, v in m.state_dict().items() if k.startswith(('predictor.', 'kernel_dist_layer.', 'init_latents'))}
    pd = kw['pred_dict']
    spec = dict(pred_type=pd.get('pred_type', 'transformer'), num_layers=pd.get('pred_num_layers', 0),
                num_heads=pd.get('pred_num_heads', 0), norm_first=pd['pred_norm_first'], rnn=pd['pred_rnn'],
                kernel_mlp=kw['slot_dict']['kernel_mlp'])
    prev, noise = TC.inputs(name)
    stochastic = kw['loss_dict']['kld_method'] != 'none'
    state = None
    for t in range(steps + 1):
        nz = noise[t] if stochastic else None
        if t == 0:
            x0 = np.repeat(w['init_latents'], B, axis=0)
            dist, init, _ = O.transition(x0, w, **dict(spec, pred_type=None), noise=nz)
        else:
            dist, init, state = O.transition(prev[t - 1], w, **spec, state=state, noise=nz)
        assert np.abs(dist - g[f'{name}.dist'][t]).max() < 2e-7 * np.abs(g[f'{name}.dist'][t]).max() + 1e-7
        assert np.abs(init - g[f'{name}.init'][t]).max() < 2e-7 * np.abs(g[f'{name}.init'][t]).max() + 1e-7
    if pd['pred_rnn']:
        assert np.abs(state[0] - g[f'{name}.h'][0]).max() < 2e-7
        assert np.abs(state[1] - g[f'{name}.c'][0]).max() < 5e-7
