#!/usr/bin/env python
"""Pipelined step time (CUDA-graph replay of 20 steps, median of 7) as a function of the Slot Attention CTA cap."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden'))
import numpy as np, torch, bench
from slotformer_b200 import engine
from slotformer_b200.base_slots.models import SlotAttention
from slotformer_b200.video_prediction.models import SlotRollouter
WL = bench.WL; dev = torch.device('cuda', 0)
sa_w, ro_w = bench.make_weights()
sa = SlotAttention(WL['C'], WL['iters'], WL['K'], WL['D'], WL['Dm']); sa.load_state_dict({k: torch.from_numpy(v) for k, v in sa_w.items()}); sa = sa.to(dev).eval()
ro = SlotRollouter(WL['K'], WL['D'], WL['T_in'], d_model=WL['d'], num_layers=WL['layers'], num_heads=WL['heads'], ffn_dim=WL['F'])
ro.load_state_dict({k: torch.from_numpy(v) for k, v in ro_w.items()}, strict=False); ro = ro.to(dev).eval()
B, T_in, T_out, K, D = WL['B'], WL['T_in'], WL['T_out'], WL['K'], WL['D']
feats = torch.randn((B * T_in, WL['N'], WL['C']), device=dev); init = torch.randn((B * T_in, K, D), device=dev)
steps = 20
with torch.no_grad():
    for cap in [int(a) for a in (sys.argv[1:] or ['84', '80', '76', '72', '68'])]:
        pipe = engine.HotPathPipeline(sa, ro, dev, clips=B)
        pipe.sa_ctas = cap
        with pipe:
            for _ in range(3): pipe.submit(feats, init, B, T_in, T_out)
        torch.cuda.synchronize()
        graph, outs = pipe.capture([(feats, init)] * steps, B, T_in, T_out)
        graph.replay(); torch.cuda.synchronize()
        reg = []
        for _ in range(7):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); graph.replay(); b.record(); torch.cuda.synchronize(); reg.append(a.elapsed_time(b) / steps)
        print(f'SA cap {cap}: pipelined {np.median(reg):.4f} ms/step (min {min(reg):.4f}, max {max(reg):.4f})', flush=True)
