#!/usr/bin/env python
"""In-kernel phase timeline (globaltimer stamps of cluster 0 / CTA 0) for both hot paths."""
import os, sys, ctypes, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden'))
import numpy as np, torch
import bench
from slotformer_b200 import engine
from slotformer_b200.base_slots.models import SlotAttention
from slotformer_b200.video_prediction.models import SlotRollouter

dev = 'cuda:0'
lib = engine.load()
WL = bench.WL
print('max co-resident clusters: C=128 cs8', lib.sfb_debug_sa_max_clusters(128, 8), 'cs16', lib.sfb_debug_sa_max_clusters(128, 16),
      '| C=192 cs16', lib.sfb_debug_sa_max_clusters(192, 16))
sa_w, ro_w = bench.make_weights()
sa = SlotAttention(WL['C'], WL['iters'], WL['K'], WL['D'], WL['Dm'])
sa.load_state_dict({k: torch.from_numpy(v) for k, v in sa_w.items()}); sa = sa.to(dev).eval()
ro = SlotRollouter(WL['K'], WL['D'], WL['T_in'], d_model=WL['d'], num_layers=WL['layers'], num_heads=WL['heads'], ffn_dim=WL['F'])
ro.load_state_dict({k: torch.from_numpy(v) for k, v in ro_w.items()}, strict=False); ro = ro.to(dev).eval()
frames = WL['B'] * WL['T_in']
feats = torch.randn((frames, WL['N'], WL['C']), device=dev)
init = torch.randn((frames, WL['K'], WL['D']), device=dev)
cap = 8192
buf = torch.zeros(cap, dtype=torch.int64, device=dev)

def timed(fn, n=5):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n

with torch.no_grad():
    for cs in (8, 16):
        sa.cluster_size = cs
        ms = timed(lambda: sa(feats, init))
        buf.zero_(); lib.sfb_debug_set_profile(buf.data_ptr(), cap)
        sa(feats, init); torch.cuda.synchronize()
        lib.sfb_debug_set_profile(None, 0)
        t = buf.cpu().numpy(); n = int((t != 0).sum()); t = t[:n].astype(np.float64)
        per_frame = 1 + 8 * WL['iters']
        nf = n // per_frame
        T = t[:nf * per_frame].reshape(nf, per_frame)
        d = np.diff(T, axis=1)
        names = ['stepD+sync', 'pass', 'tree-reduce', 'E1 sync', 'owner+E2 sync', 'stepA+E3', 'stepB+E4', 'stepC']
        print(f'--- SA cluster_size={cs}: {ms*1e3:.0f} us/launch, {nf} frames on cluster 0, frame period {np.diff(T[:,0]).mean()/1e3:.2f} us')
        for it in range(WL['iters']):
            print(f'  iter {it}: ' + ', '.join(f'{names[j]} {d[1:, it*8+j].mean()/1e3:.2f}' for j in range(8)))
        if nf > 1:
            print(f'  gap lastC->next frame start: {((T[1:,0]-T[:-1,-1]).mean())/1e3:.2f} us')
    sa.cluster_size = 0
    slots = sa(feats, init).view(WL['B'], WL['T_in'], WL['K'], WL['D']).contiguous()
    ms = timed(lambda: ro(slots, WL['T_out']))
    buf.zero_(); lib.sfb_debug_set_profile(buf.data_ptr(), cap)
    ro(slots, WL['T_out']); torch.cuda.synchronize()
    lib.sfb_debug_set_profile(None, 0)
    t = buf.cpu().numpy(); n = int((t != 0).sum()); t = t[:n].astype(np.float64)
    per_layer = 7 + 2 * 1   # LN1, qkv, attn, out, LN2, (ffn1, ffn2) x2 chunks -> stamps: LN1,qkv,attn,outproj,LN2,ffn1a,(ffn2a) ...
    print(f'--- rollout: {ms*1e3:.0f} us/launch, {n} stamps')
    d = np.diff(t)
    per_step = n // WL['T_out']
    print('  per-step stamps', per_step, ' step period us', (t[per_step] - t[0]) / 1e3)
    print('  first step deltas (us):', np.round(d[:per_step] / 1e3, 2).tolist())
