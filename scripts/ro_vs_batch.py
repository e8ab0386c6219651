#!/usr/bin/env python
"""Rollout (bench config, one CTA per clip) timed alone as a function of the clip count: every CTA re-reads all weight tiles
from L2, so the slope over B is the kernel's sensitivity to L2 traffic (the pipelined step is rollout-bound)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden'))
import numpy as np, torch, bench
from slotformer_b200.video_prediction.models import SlotRollouter
WL = bench.WL; dev = torch.device('cuda', 0)
_, ro_w = bench.make_weights()
ro = SlotRollouter(WL['K'], WL['D'], WL['T_in'], d_model=WL['d'], num_layers=WL['layers'], num_heads=WL['heads'], ffn_dim=WL['F'])
ro.load_state_dict({k: torch.from_numpy(v) for k, v in ro_w.items()}, strict=False); ro = ro.to(dev).eval()
with torch.no_grad():
    for B in (8, 32, 64, 96, 128, 148):
        x = torch.randn((B, WL['T_in'], WL['K'], WL['D']), device=dev)
        for _ in range(3): ro(x, WL['T_out'])
        torch.cuda.synchronize()
        ts = []
        for _ in range(10):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); ro(x, WL['T_out']); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b) * 1e3)
        print(f'B={B:4d}: {np.median(ts):.1f} us', flush=True)
