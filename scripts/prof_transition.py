#!/usr/bin/env python
"""Per-step timeline of the fused slot transition (CTA 0, globaltimer stamps; debug library)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests')); sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden'))
import numpy as np, torch
import transition_cases as TC
from slotformer_b200 import engine
from slotformer_b200.base_slots.models import StoSAVi
dev = 'cuda:0'; lib = engine.use_debug_library()
m = TC.build(StoSAVi, 'tr_obj3d').to(dev)
spec = m._transition_spec()
eng = engine.TransitionEngine()
NAMES = ['end', 'LOAD', 'STORE', 'LN', 'LINEAR', 'ATTN', 'LSTM', 'SAMPLE']
# the program of tr_obj3d (capi.cu tr_program): for the printout only
PROG = ['LOAD', 'LOAD h', 'LOAD c'] + ['LN', 'LINEAR qkv', 'ATTN', 'LINEAR wo', 'LN', 'LINEAR w1', 'LINEAR w2'] * 2 + \
       ['LINEAR gates', 'LSTM', 'LINEAR out', 'LINEAR kd0', 'LN', 'LINEAR kd3', 'STORE', 'SAMPLE']
with torch.no_grad():
    for B in [int(a) for a in sys.argv[1:]] or [4, 64, 148]:
        prev = torch.randn(B, 6, 128, device=dev)
        h = torch.randn(B * 6, 256, device=dev); c = torch.randn(B * 6, 256, device=dev)
        a = torch.randn(4096, 4096, device=dev)
        for _ in range(300): a @ a                     # SM clock up
        for _ in range(3): eng.forward(spec, prev, True, B, (h, c))
        buf = torch.zeros(128, dtype=torch.int64, device=dev)
        torch.cuda.synchronize()
        lib.sfb_debug_set_profile(buf.data_ptr(), 128)
        eng.forward(spec, prev, True, B, (h, c)); torch.cuda.synchronize()
        lib.sfb_debug_set_profile(None, 0)
        t = buf.cpu().numpy(); n = int((t != 0).sum()); t = t[:n].astype(np.float64) / 1965.0      # clock64 at 1965 MHz -> us
        d = np.diff(t)
        print(f'B={B}: {n - 1} steps, {t[-1] - t[0]:.1f} us in the kernel')
        print('  ' + ' | '.join(f'{PROG[i] if i < len(PROG) else i} {d[i]:.2f}' for i in range(len(d))))
