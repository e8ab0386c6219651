#!/usr/bin/env python
"""A few back-to-back calls of the two hot operators on the bench workload (for ncu captures).
SA_CTAS=<n> caps the Slot Attention grid like the batch pipeline does."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden'))
import torch, bench
from slotformer_b200 import engine
from slotformer_b200.base_slots.models import SlotAttention
from slotformer_b200.video_prediction.models import SlotRollouter
WL = bench.WL; dev = 'cuda:0'
sa_w, ro_w = bench.make_weights()
sa = SlotAttention(WL['C'], WL['iters'], WL['K'], WL['D'], WL['Dm']); sa.load_state_dict({k: torch.from_numpy(v) for k, v in sa_w.items()}); sa = sa.to(dev).eval()
ro = SlotRollouter(WL['K'], WL['D'], WL['T_in'], d_model=WL['d'], num_layers=WL['layers'], num_heads=WL['heads'], ffn_dim=WL['F'])
ro.load_state_dict({k: torch.from_numpy(v) for k, v in ro_w.items()}, strict=False); ro = ro.to(dev).eval()
feats = torch.randn((384, 4096, 128), device=dev); init = torch.randn((384, 6, 128), device=dev)
sa.max_ctas = int(os.environ.get('SA_CTAS', '0'))
if os.environ.get('SA_NO_TC'):
    sa.engine_flags = engine.SFB_SA_NO_TCGEN05
with torch.no_grad():
    for _ in range(int(os.environ.get('REPS', '3'))):
        s = sa(feats, init)
        if not os.environ.get('SA_ONLY'):
            ro(s.view(64, 6, 6, 128), 10)
    if not os.environ.get('SA_ONLY'):
        x = torch.randn((640, 6, 4, 128, 128), device=dev)
        engine.decode_combine(x, want_seg=True)
torch.cuda.synchronize()
