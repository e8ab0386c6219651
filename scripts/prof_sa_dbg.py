#!/usr/bin/env python
"""Slot Attention time with parts of the pass kernel disabled (SFB_DBG: 1 no proxy fence, 2 no x^ store,
4 no LN / MMA work = pure TMA streaming) at several CTA limits."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden'))
import torch, bench
from slotformer_b200 import engine
engine.use_debug_library()   # the SFB_DBG switches exist only in the -DSFB_DEBUG build
from slotformer_b200.base_slots.models import SlotAttention
WL = bench.WL; dev = 'cuda:0'
sa_w, _ = bench.make_weights()
sa = SlotAttention(WL['C'], WL['iters'], WL['K'], WL['D'], WL['Dm']); sa.load_state_dict({k: torch.from_numpy(v) for k, v in sa_w.items()}); sa = sa.to(dev).eval()
feats = torch.randn((384, 4096, 128), device=dev); init = torch.randn((384, 6, 128), device=dev)
with torch.no_grad():
    for lim in (0, 84):
        sa.max_ctas = lim
        for dbg in (0,):
            os.environ['SFB_DBG'] = str(dbg)
            for _ in range(3): sa(feats, init)
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(10): sa(feats, init)
            b.record(); torch.cuda.synchronize()
            print(f'cta limit {lim} dbg {dbg}: SA {a.elapsed_time(b) / 10 * 1e3:.0f} us', flush=True)
