#!/usr/bin/env python
"""Timing of the fused-encoder-tail route on the bench workload (384 frames): sfb_enc_tail_forward (CNN output ->
operand tiles) and Slot Attention on the tiles, next to Slot Attention on the fp32 feature grid."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests')); sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden'))
import numpy as np, torch
import bench, wrapper_cases as W
from helpers import sa_module
from slotformer_b200 import engine
from slotformer_b200.base_slots.models import StoSAVi
dev = 'cuda:0'
frames = int(os.environ.get('FRAMES', '384'))
m = W.build_savi(StoSAVi).to(dev).eval()
named = dict(m.named_parameters())
wts = {k: named[k].detach() for k in engine.ENC_TAIL_KEYS}
sa_w, _ = bench.make_weights()
sa = sa_module(dict(C=128, iters=2, K=6, D=128, Dm=256, mask=False), sa_w, dev)
gen = torch.Generator(device=dev).manual_seed(3)
cnn = torch.randn((frames, 64, 64, 64), device=dev, generator=gen)
init = torch.randn((frames, 6, 128), device=dev, generator=gen)
tail = engine.EncoderTailEngine()

def timed(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    return float(np.median(ts))

with torch.no_grad():
    for ctas in (0, 84):
        sa.max_ctas = ctas
        tiles = tail.forward(cnn, wts, 128, max_ctas=ctas)
        t_tail = timed(lambda: tail.forward(cnn, wts, 128, max_ctas=ctas))
        t_sa_tiles = timed(lambda: sa(tiles, init))
        feats = tiles.to_dense().contiguous()
        t_sa_grid = timed(lambda: sa(feats, init))
        a, b = sa(tiles, init), sa(feats, init)
        cnn_cl = cnn.contiguous(memory_format=torch.channels_last)
        t_tail_cl = timed(lambda: tail.forward(cnn_cl, wts, 128, max_ctas=ctas))
        print(f'ctas {ctas or 148}: enc_tail on channels-last input (SFB_ET_NHWC) {t_tail_cl:.0f} us', flush=True)
        print(f'ctas {ctas or 148}: enc_tail {t_tail:.0f} us ({frames * 4096 * (64 * 4 + 128 * 2) / t_tail / 1e6:.2f} TB/s in+out), '
              f'SA on tiles {t_sa_tiles:.0f} us, SA on the fp32 grid {t_sa_grid:.0f} us, '
              f'tiles vs grid slots rel {float((a - b).abs().max() / b.abs().max()):.2e}', flush=True)
