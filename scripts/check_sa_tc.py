#!/usr/bin/env python
"""tcgen05 Slot Attention passes against the mma.sync passes and the oracle on a few shapes, then timings at the
bench workload (384 frames x 4096 x 128) with every SM and under the batch pipeline's 84-CTA cap."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden')); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np, torch, bench, cases
from helpers import sa_module, rel_max
from slotformer_b200 import engine
from oracle import slot_oracle as O
dev = 'cuda:0'
for (B, K, N, bf16, mask) in [(1, 4, 128, False, False), (2, 6, 256, False, True), (3, 6, 1000, False, True), (2, 7, 520, False, False),
                              (5, 8, 4096, True, True), (1, 1, 64, False, False), (40, 6, 4096, False, False)]:
    c = dict(B=B, N=N, C=128, D=128, Dm=256, K=K, iters=2, mask=mask, seed=60 + K)
    w = cases.make_sa_weights(128, 128, 256, c['seed'])
    feats, slots = cases.make_sa_inputs(B, N, 128, 128, K, c['seed'])
    feats = feats + 0.5
    m = sa_module(c, w, dev)
    f = torch.from_numpy(feats).to(dev); s0 = torch.from_numpy(slots).to(dev)
    if bf16:
        f = f.to(torch.bfloat16); feats = f.float().cpu().numpy()
    with torch.no_grad():
        m.engine_flags = engine.SFB_SA_NO_TCGEN05
        a = m(f, s0)
        m.engine_flags = 0
        b = m(f, s0); b2 = m(f, s0)
    torch.cuda.synchronize()
    if mask:
        (a, am), (b, bm), (b2, _) = a, b, b2
    if B <= 5:
        ref = O.slot_attention(feats, slots, w, 2, return_mask=mask)
        if mask: ref, rm = ref
        msg = f'oracle: mma {rel_max(a.cpu().numpy(), ref):.2e} tc {rel_max(b.cpu().numpy(), ref):.2e}'
        if mask: msg += f' mask: mma {np.abs(am.cpu().numpy() - rm).max():.2e} tc {np.abs(bm.cpu().numpy() - rm).max():.2e}'
    else:
        msg = ''
    print(f'B={B} K={K} N={N} bf16={bf16}: tc vs mma {rel_max(b.cpu().numpy(), a.cpu().numpy()):.2e} repeat-equal {torch.equal(b, b2)} finite {bool(torch.isfinite(b).all())} {msg}', flush=True)

WL = bench.WL
sa_w, _ = bench.make_weights()
c = dict(B=384, N=4096, C=128, D=128, Dm=256, K=6, iters=2, mask=False)
sa = sa_module(c, sa_w, dev)
feats = torch.randn((384, 4096, 128), device=dev); init = torch.randn((384, 6, 128), device=dev)
with torch.no_grad():
    for flags, name in ((engine.SFB_SA_NO_TCGEN05, 'mma.sync'), (0, 'tcgen05')):
        for lim in (0, 84):
            sa.engine_flags = flags; sa.max_ctas = lim
            for _ in range(3): sa(feats, init)
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(10): sa(feats, init)
            b.record(); torch.cuda.synchronize()
            print(f'{name} cta limit {lim}: SA {a.elapsed_time(b) / 10 * 1e3:.0f} us', flush=True)
