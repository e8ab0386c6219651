#!/usr/bin/env python
"""Per-phase rollout timeline (CTA 0 stamps) with parts of the kernel disabled via SFB_DBG."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden'))
import numpy as np, torch, bench
from slotformer_b200 import engine
from slotformer_b200.video_prediction.models import SlotRollouter
dev = 'cuda:0'; WL = dict(bench.WL); lib = engine.use_debug_library()   # -DSFB_DEBUG build: timeline hook + SFB_DBG switches
CASE = os.environ.get('RO_CASE')          # e.g. RO_CASE=ro_cfg3: a tests/golden/cases.py rollout case instead of the bench workload
if CASE:
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    import cases
    from helpers import ro_module
    c, w, hist = cases.ro_case(CASE)
    g = np.load(os.path.join(ROOT, 'tests', 'golden', CASE + '.npz'))
    ro = ro_module(c, w, dev, enc_t_pe=g['enc_t_pe'])
    WL.update(B=int(os.environ.get('RO_B', 32)), T_in=c['T_h'], K=c['K'], D=c['Ds'], layers=c['layers'], T_out=min(c['pred_len'], 10))
else:
    _, ro_w = bench.make_weights()
    ro = SlotRollouter(WL['K'], WL['D'], WL['T_in'], d_model=WL['d'], num_layers=WL['layers'], num_heads=WL['heads'], ffn_dim=WL['F'])
    ro.load_state_dict({k: torch.from_numpy(v) for k, v in ro_w.items()}, strict=False); ro = ro.to(dev).eval()
x = torch.randn((WL['B'], WL['T_in'], WL['K'], WL['D']), device=dev)
cap = 8192
buf = torch.zeros(cap, dtype=torch.int64, device=dev)
names = ['in_proj', 'LN1', 'qkv', 'attn', 'outproj', 'LN2', 'ffn', 'x7', 'x8']
dbgs = [int(a) for a in sys.argv[1:]] or [0, 127, 103, 100, 4, 2]
with torch.no_grad():
    for dbg in dbgs:
        os.environ['SFB_DBG'] = str(dbg)
        for _ in range(2): ro(x, WL['T_out'])
        torch.cuda.synchronize()
        buf.zero_(); lib.sfb_debug_set_profile(buf.data_ptr(), cap)
        ro(x, WL['T_out']); torch.cuda.synchronize()
        lib.sfb_debug_set_profile(None, 0)
        t = buf.cpu().numpy()[:2048]; n = int((t != 0).sum()); t = t[:n].astype(np.float64) / 1e3
        per_step = n // WL['T_out']
        T = t[:per_step * WL['T_out']].reshape(WL['T_out'], per_step)
        d = np.diff(T, axis=1)            # [steps, per_step-1]
        nl = WL['layers']; ppl = (per_step - 2) // nl
        lay = d[:, 1:1 + nl * ppl].reshape(WL['T_out'], nl, ppl).mean(axis=(0, 1))
        tail = (T[1:, 0] - T[:-1, -1]).mean()
        print(f'dbg={dbg}: step {np.diff(T[:, 0]).mean():.2f} us | in_proj {d[:, 0].mean():.2f} | ' +
              ' '.join(f'{names[1 + j]}={lay[j]:.2f}' for j in range(ppl)) + f' | layer {lay.sum():.2f} | tail {tail:.2f}', flush=True)
