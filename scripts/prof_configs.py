#!/usr/bin/env python
"""Both hot operators on every BASELINE.json config shape (per-GPU batch): time, algorithmic GB/s / TFLOP/s."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests')); sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden'))
import numpy as np, torch
import cases
from helpers import ro_module, sa_module
dev = 'cuda:0'

def timed(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n

# (name, frames per GPU batch, bf16 features)
SA = [('sa_cfg1', 4 * 6, False), ('sa_cfg2', 64 * 6, False), ('sa_cfg3', 32 * 6, True), ('sa_cfg4', 16 * 4, False)]
RO = [('ro_cfg2', 64), ('ro_cfg3', 32), ('ro_cfg4', 16), ('ro_cfg5', 256), ('ro_physion', 16)]
with torch.no_grad():
    for name, frames, bf16 in ([] if os.environ.get('SKIP_SA') else SA):
        c, w, _, _ = cases.sa_case(name)
        m = sa_module(c, w, dev, mask=c['mask'])
        f = torch.randn((frames, c['N'], c['C']), device=dev)
        if bf16: f = f.to(torch.bfloat16)
        s0 = torch.randn((frames, c['K'], c['D']), device=dev)
        ms = timed(lambda: m(f, s0))
        nbytes = frames * (c['N'] * c['C'] * (2 if bf16 else 4) + 2 * c['K'] * c['D'] * 4 + (c['K'] * c['N'] * 4 if c['mask'] else 0))
        print(f'{name}: {frames} frames N={c["N"]} C={c["C"]} K={c["K"]} iters={c["iters"]} {"bf16" if bf16 else "fp32"}: '
              f'{ms*1e3:.0f} us, {nbytes/ms/1e6:.0f} GB/s algorithmic', flush=True)
    for name, B in RO:
        c, w, hist = cases.ro_case(name)
        g = np.load(os.path.join(ROOT, 'tests', 'golden', name + '.npz'))
        m = ro_module(c, w, dev, enc_t_pe=g['enc_t_pe'])
        x = torch.randn((B,) + hist.shape[1:], device=dev)
        ms = timed(lambda: m(x, c['pred_len']), n=5)
        K, Ds, d, F, Ly = c['K'], c['Ds'], c['d'], c['F'], c['layers']
        fl = 0
        for s in range(c['pred_len']):
            L = c['T_h'] * K if c['mode'] != 'grow' else min(K * (1 + s), c['cond_len'] * K)
            fl += 2 * L * Ds * d + Ly * (8 * L * d * d + 4 * L * L * d + 4 * L * d * F) + 2 * K * d * Ds
        print(f'{name}: B={B} K={K} d={d} F={F} layers={Ly} steps={c["pred_len"]} mode={c["mode"]}: {ms*1e3:.0f} us, '
              f'{fl*B/ms/1e9:.1f} TFLOP/s, {ms*1e3/c["pred_len"]/Ly:.1f} us per layer-step', flush=True)
