#!/usr/bin/env python
"""Small rollout cases for compute-sanitizer (synccheck / racecheck on the tcgen05 rollout engine)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests')); sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden'))
import numpy as np, torch
import cases
from helpers import golden, rel_max, ro_module
dev = 'cuda:0'
with torch.no_grad():
    for name in sys.argv[1:] or ['ro_tiny', 'ro_cfg2']:
        c, w, hist = cases.ro_case(name)
        g = golden(name)
        m = ro_module(c, w, dev, enc_t_pe=g['enc_t_pe'])
        out = m(torch.from_numpy(hist).to(dev), min(c['pred_len'], 3))
        torch.cuda.synchronize()
        print(name, f'{rel_max(out.cpu().numpy(), g["pred_f64"][:, :out.shape[1]]):.2e}', flush=True)
print('done')
