#!/usr/bin/env python
"""SAVi frame loop (predictor -> distribution head -> Slot Attention per frame): eager launches vs one CUDA-graph replay."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden'))
import torch, wrapper_cases as W
from slotformer_b200.base_slots.models import StoSAVi
m = W.build_savi(StoSAVi).cuda()
img = torch.cat([W.savi_input()] * 2, dim=0).cuda()
def t(fn, n=10):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3
with torch.no_grad():
    feats = m._get_encoder_out(img.flatten(0, 1)).unflatten(0, (4, 3)).contiguous()
    print('encoder only: %.2f ms' % t(lambda: m._get_encoder_out(img.flatten(0, 1))))
    def eager(): m.predictor.reset(); return m._frame_loop(feats, None)
    def graphed(): m.predictor.reset(); return m._frame_loop_graphed(feats, None)
    print('frame loop eager: %.2f ms' % t(eager))
    print('frame loop graphed: %.2f ms' % t(graphed))
    ent = list(m._loop_graphs.values())[-1]
    print('replay only: %.2f ms' % t(lambda: ent[0].replay()))
    print('key only: %.3f ms' % t(lambda: m._graph_key(feats, None, False)))
