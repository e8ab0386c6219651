#!/usr/bin/env python
"""Rollout kernel phase timeline (CTA 0) on the bench workload."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden'))
import numpy as np, torch, bench
from slotformer_b200 import engine
from slotformer_b200.video_prediction.models import SlotRollouter
dev = 'cuda:0'; WL = bench.WL; lib = engine.load()
_, ro_w = bench.make_weights()
ro = SlotRollouter(WL['K'], WL['D'], WL['T_in'], d_model=WL['d'], num_layers=WL['layers'], num_heads=WL['heads'], ffn_dim=WL['F'])
ro.load_state_dict({k: torch.from_numpy(v) for k, v in ro_w.items()}, strict=False); ro = ro.to(dev).eval()
x = torch.randn((WL['B'], WL['T_in'], WL['K'], WL['D']), device=dev)
cap = 4096; buf = torch.zeros(cap, dtype=torch.int64, device=dev)
with torch.no_grad():
    for _ in range(3): ro(x, WL['T_out'])
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10): ro(x, WL['T_out'])
    b.record(); torch.cuda.synchronize()
    print(f'rollout: {a.elapsed_time(b)/10*1e3:.0f} us/launch')
    lib.sfb_debug_set_profile(buf.data_ptr(), cap); ro(x, WL['T_out']); torch.cuda.synchronize(); lib.sfb_debug_set_profile(None, 0)
t = buf.cpu().numpy(); n = int((t != 0).sum()); t = t[:n].astype(np.float64)
per_step = n // WL['T_out']
T = t[:per_step * WL['T_out']].reshape(WL['T_out'], per_step)
d = np.diff(T, axis=1).mean(0) / 1e3
d = np.diff(T, axis=1).mean(0) / 1e3
print('stamps/step', per_step, 'step period us', np.diff(T[:, 0]).mean() / 1e3)
print('deltas (us):', np.round(d, 2).tolist())
print('tail us', (np.diff(T[:, 0]).mean() - (T[:, -1] - T[:, 0]).mean()) / 1e3)
