#!/usr/bin/env python
"""Slot Attention error against the reference goldens for every case (max|out-ref| / max|ref|; tolerance 1e-3)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests')); sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden'))
import numpy as np, torch
import cases
from helpers import golden, rel_max, sa_module
for name in cases.SA_CASES:
    c, w, feats, slots = cases.sa_case(name)
    g = golden(name)
    m = sa_module(c, w, 'cuda:0')
    with torch.no_grad():
        out = m(torch.from_numpy(feats).cuda(), torch.from_numpy(slots).cuda())
    mask_err = ''
    if c['mask']:
        out, mask = out
        mask_err = f'  mask abs {np.abs(mask.cpu().numpy() - g["mask_f64"]).max():.2e}'
    print(f'{name}: slots rel {rel_max(out.cpu().numpy(), g["slots_f64"]):.2e}{mask_err}', flush=True)
