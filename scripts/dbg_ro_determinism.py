#!/usr/bin/env python
"""Repeatability of the rollout on a BASELINE config: which clips / steps differ between runs?"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests')); sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden'))
import numpy as np, torch
import cases
from helpers import ro_module
name = sys.argv[1] if len(sys.argv) > 1 else 'ro_cfg5'
B = int(sys.argv[2]) if len(sys.argv) > 2 else 256
c, w, hist = cases.ro_case(name)
g = np.load(os.path.join(ROOT, 'tests', 'golden', name + '.npz'))
m = ro_module(c, w, 'cuda:0', enc_t_pe=g['enc_t_pe'])
gen = torch.Generator(device='cuda:0').manual_seed(11)
x = torch.randn((B,) + hist.shape[1:], device='cuda:0', generator=gen)
with torch.no_grad():
    a = m(x, c['pred_len'])
    for rep in range(6):
        b = m(x, c['pred_len'])
        diff = (a - b).abs().amax(dim=(2, 3))           # [B, steps]
        bad = (diff > 0).nonzero()
        print(f'rep {rep}: max diff {diff.max().item():.3e}, {len(bad)} (clip, step) cells differ; first:', bad[:6].tolist(), flush=True)
