#!/usr/bin/env python
"""Slot transition (predictor -> kernel_dist_layer -> sample): fused kernel vs the stock PyTorch modules, and the
SAVi frame loop (config-1-like extraction: B clips, T frames) with / without it."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests')); sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden'))
import torch
import transition_cases as TC
from slotformer_b200 import engine
from slotformer_b200.base_slots.models import StoSAVi
dev = 'cuda:0'

def burn(ms=300):
    # bring the SM clock up before a microsecond-scale measurement (an idle GPU sits at a low clock)
    a = torch.randn(4096, 4096, device=dev)
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    while True:
        for _ in range(10): a @ a
        t1.record(); torch.cuda.synchronize()
        if t0.elapsed_time(t1) > ms: break

def timed(fn, n=200):
    burn(100)
    for _ in range(5): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3

m = TC.build(StoSAVi, 'tr_obj3d').to(dev)
spec = m._transition_spec()
eng = engine.TransitionEngine()
with torch.no_grad():
    for B in (1, 4, 18, 37, 64, 148, 384):
        prev = torch.randn(B, 6, 128, device=dev)
        h = torch.randn(B * 6, 256, device=dev); c = torch.randn(B * 6, 256, device=dev)
        def stock():
            m.predictor.hidden_state = (h[None], c[None])
            return m.kernel_dist_layer(m.predictor(prev))
        g = torch.cuda.CUDAGraph()
        m.predictor.rnn.flatten_parameters()
        stock(); torch.cuda.synchronize()
        with torch.cuda.graph(g):
            stock()
        t_f = timed(lambda: eng.forward(spec, prev, True, B, (h, c)))
        gf = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gf):
            eng.forward(spec, prev, True, B, (h, c))
        print(f'B={B:4d}: fused {t_f:7.1f} us (graph replay {timed(gf.replay):6.1f} us) | stock eager {timed(stock):7.1f} us, '
              f'stock graph replay {timed(g.replay):6.1f} us', flush=True)
    # SAVi extraction loop, config-1-like: B=4 clips, T=6 frames, 64x64 images, K=6
    img = torch.rand(4, 6, 3, 64, 64, device=dev) * 2 - 1
    for fused in (False, True):
        for graph in (False, True):
            m.fuse_transition, m.use_cuda_graph = fused, graph
            m.__dict__.pop('_loop_graphs', None)
            def run():
                m._reset_rnn()
                return m.encode(img)
            print(f'StoSAVi.encode B=4 T=6: fused_transition={fused} cuda_graph={graph}: {timed(run, 20):8.1f} us per call', flush=True)
