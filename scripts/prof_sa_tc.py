#!/usr/bin/env python
"""Per-role clock64 timeline of CTA 0 of the tcgen05 Slot Attention passes (debug build): where does a tile's time go?
roles: 0 MMA issuer, 1 softmax warp 0, 2 / 3 LN warp 0 / 4 (first pass)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden')); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np, torch, bench
from slotformer_b200 import engine
lib = engine.use_debug_library()
from helpers import sa_module
dev = 'cuda:0'
sa_w, _ = bench.make_weights()
TAGS = {0: {1: 'wait-tile', 2: 'got-tile(L issue)', 3: 'wait-P', 4: 'got-P(agg issue)'},
        1: {1: 'wait-logits', 2: 'got-logits', 3: 'ldtm-done', 4: 'P-stored', 5: 'P-arrived', 6: 'wait-acc', 7: 'got-acc'},
        2: {1: 'wait-stage', 2: 'got-stage', 3: 'wait-tfree', 4: 'got-tfree', 5: 'store-read-done', 6: 'rows-written', 7: 'fenced', 8: 'arrived'}}
TAGS[3] = TAGS[2]
TAGS[4] = {3: 'agg wait-P', 4: 'agg got-P (issue)'}
cap = 8 * 512
buf = torch.zeros(cap, dtype=torch.int64, device=dev)
for iters, name in ((1, 'FIRST pass only (no x^ store)'), (2, 'FIRST + NEXT')):
    c = dict(B=384, N=4096, C=128, D=128, Dm=256, K=6, iters=iters, mask=False)
    sa = sa_module(c, sa_w, dev)
    sa.max_ctas = int(os.environ.get('SA_CTAS', '0'))      # e.g. SA_CTAS=84: the batch pipeline's cap
    feats = torch.randn((384, 4096, 128), device=dev); init = torch.randn((384, 6, 128), device=dev)
    with torch.no_grad():
        for _ in range(2): sa(feats, init)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); sa(feats, init); b.record(); torch.cuda.synchronize()
        print(f'=== {name}: SA call {a.elapsed_time(b) * 1e3:.0f} us')
        buf.zero_(); lib.sfb_debug_set_profile(buf.data_ptr(), cap)
        sa(feats, init); torch.cuda.synchronize()
        lib.sfb_debug_set_profile(None, 0)
    t = buf.cpu().numpy().astype(np.uint64)
    ev = []
    for role in range(5):
        seg = t[role * 512:(role + 1) * 512]; seg = seg[seg != 0]
        for v in seg:
            ev.append((int(v & np.uint64(0xFFFFFFFFFF)), role, int(v >> np.uint64(56)), int((v >> np.uint64(40)) & np.uint64(0xffff))))
    ev.sort()
    # the LAST kernel that wrote the buffer wins (iters=2: the NEXT pass overwrote roles 0 / 1); print tiles 8..13
    if not ev: continue
    # LayerNorm warp 0 (role 2): average clocks per phase over its sub-tiles
    r2 = [(clk, tag) for clk, role, tag, tile in ev if role == 2]
    if len(r2) > 16:
        names = TAGS[2]; acc = {}
        for (c0, t0_), (c1, t1_) in zip(r2[:-1], r2[1:]):
            acc.setdefault(f'{names.get(t0_, t0_)} -> {names.get(t1_, t1_)}', []).append(c1 - c0)
        print('   LN warp 0, mean clocks per transition: ' + '; '.join(f'{k}: {np.mean(v):.0f} (n={len(v)})' for k, v in acc.items()))
    t0 = ev[0][0]
    for clk, role, tag, tile in ev:
        if 8 <= tile <= 12:
            print(f'{clk - t0:9d}  role {role}  tile {tile:3d}  {TAGS[role].get(tag, tag)}')
