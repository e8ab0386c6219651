#!/usr/bin/env python
"""Times sfb_sa_forward on the bench workload (event-timed), for SFB_DBG variants / chunk sizes."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden'))
import torch, bench
from slotformer_b200.base_slots.models import SlotAttention
dev = 'cuda:0'; WL = bench.WL
sa_w, _ = bench.make_weights()
sa = SlotAttention(WL['C'], WL['iters'], WL['K'], WL['D'], WL['Dm'])
sa.load_state_dict({k: torch.from_numpy(v) for k, v in sa_w.items()}); sa = sa.to(dev).eval()
frames = WL['B'] * WL['T_in']
feats = torch.randn((frames, WL['N'], WL['C']), device=dev)
init = torch.randn((frames, WL['K'], WL['D']), device=dev)
def timed(fn, n=5):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3
with torch.no_grad():
    if os.environ.get('BF16'):
        feats = feats.to(torch.bfloat16)
    for chunk in [int(x) for x in os.environ.get('CHUNKS', '0,384').split(',')]:
        sa.chunk_frames = chunk
        for it in (1, 2):
            sa.num_iterations = it
            print(f'SFB_DBG={os.environ.get("SFB_DBG","0")} chunk={chunk} iters={it}: {timed(lambda: sa(feats, init)):.0f} us')
