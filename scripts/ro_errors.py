#!/usr/bin/env python
"""Rollout error against the reference goldens for every case: first step and worst free-running step
(relative Frobenius error and max-norm error; bounds in tests/test_gpu_parity.py)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests')); sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden'))
import numpy as np, torch
import cases
from helpers import golden, rel_l2, rel_max, ro_module
for name in cases.RO_CASES:
    c, w, hist = cases.ro_case(name)
    g = golden(name)
    m = ro_module(c, w, 'cuda:0', enc_t_pe=g['enc_t_pe'])
    with torch.no_grad():
        out = m(torch.from_numpy(hist).cuda(), c['pred_len']).cpu().numpy()
    ref = g['pred_f64']
    e2 = [rel_l2(out[:, s], ref[:, s]) for s in range(ref.shape[1])]
    em = [rel_max(out[:, s], ref[:, s]) for s in range(ref.shape[1])]
    print(f'{name}: step 0 l2 {e2[0]:.2e} max {em[0]:.2e}; worst ratio to bound (1+0.05 s): '
          f'l2 {max(e / (1 + 0.05 * s) for s, e in enumerate(e2)):.2e} max {max(e / (1 + 0.05 * s) for s, e in enumerate(em)):.2e}', flush=True)
