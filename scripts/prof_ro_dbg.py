#!/usr/bin/env python
"""Rollout time with parts of the kernel disabled (SFB_DBG bits): what bounds a phase?
bits: 1 no weight TMA, 2 no MMA, 4 no epilogue, 8 early stage release, 16 no proxy fence, 32 no attention, 64 no LN."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden'))
import torch, bench
from slotformer_b200.video_prediction.models import SlotRollouter
from slotformer_b200 import engine
engine.use_debug_library()   # the SFB_DBG switches exist only in the -DSFB_DEBUG build
dev = 'cuda:0'; WL = bench.WL
_, ro_w = bench.make_weights()
ro = SlotRollouter(WL['K'], WL['D'], WL['T_in'], d_model=WL['d'], num_layers=WL['layers'], num_heads=WL['heads'], ffn_dim=WL['F'])
ro.load_state_dict({k: torch.from_numpy(v) for k, v in ro_w.items()}, strict=False); ro = ro.to(dev).eval()
x = torch.randn((WL['B'], WL['T_in'], WL['K'], WL['D']), device=dev)
with torch.no_grad():
    for dbg in (0, 1, 2, 3, 4, 7, 32, 64, 96, 100, 103, 127):
        os.environ['SFB_DBG'] = str(dbg)
        for _ in range(2): ro(x, WL['T_out'])
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(5): ro(x, WL['T_out'])
        b.record(); torch.cuda.synchronize()
        print(f'dbg={dbg}: {a.elapsed_time(b)/5*1e3:.0f} us/launch', flush=True)
