"""Shared helpers for the parity tests (build modules from the seeded golden cases)."""
import os

import numpy as np
import torch

import cases

GOLD = os.path.join(os.path.dirname(__file__), 'golden')


def golden(name):
    return np.load(os.path.join(GOLD, f'{name}.npz'))


def rel_max(a, b):
    """max |a-b| / max |b|  -- the tolerance metric used throughout."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-30))


def rel_l2(a, b):
    """||a-b||_F / ||b||_F -- the relative error of the whole tensor."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-30))


def sa_module(c, w, device, mask=None):
    from slotformer_b200.base_slots.models import SlotAttention, SlotAttentionWMask
    cls = SlotAttentionWMask if (c['mask'] if mask is None else mask) else SlotAttention
    m = cls(in_features=c['C'], num_iterations=c['iters'], num_slots=c['K'], slot_size=c['D'],
            mlp_hidden_size=c['Dm'])
    m.load_state_dict({k: torch.from_numpy(v) for k, v in w.items()}, strict=True)
    return m.to(device).eval()


def ro_module(c, w, device, enc_t_pe=None):
    from slotformer_b200.video_prediction.models import SlotRollouter, SingleStepSlotRollouter
    kw = dict(num_slots=c['K'], slot_size=c['Ds'], history_len=c['T_h'], d_model=c['d'],
              num_layers=c['layers'], num_heads=c['heads'], ffn_dim=c['F'])
    m = SingleStepSlotRollouter(cond_len=c['cond_len'], **kw) if c['mode'] == 'grow' \
        else SlotRollouter(**kw)
    sd = {k: torch.from_numpy(v) for k, v in w.items()}
    if enc_t_pe is not None:
        sd['enc_t_pe'] = torch.from_numpy(np.asarray(enc_t_pe))
    missing, unexpected = m.load_state_dict(sd, strict=False)
    assert not unexpected and set(missing) <= {'enc_t_pe'}
    return m.to(device).eval()
