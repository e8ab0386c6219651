#!/usr/bin/env python
"""Fine per-role trace (clock64) of one rollout layer: producer / MMA thread / compute thread 0."""
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
dbgs = [int(a) for a in sys.argv[1:]] or [0]
TAGS = {'P': {1: 'wait-empty', 2: 'got-empty'}, 'M': {1: 'wait-full', 2: 'got-full', 3: 'issued+commit', 4: 'wait-rdy', 5: 'got-rdy'},
        'C': {1: 'wait-acc', 2: 'got-acc', 3: 'rdy-arrive', 20: 'ldtm-issued', 21: 'ldtm-waited', 22: 'stored', 23: 'tc-fence', 24: 'proxy-fence'}}
COMMON = {30: 'LN2-begin', 31: 'LN2-done', 32: 'LN2-synced', 9: 'LAYER-START', 10: 'LAYER-END', 11: 'ph:LN1', 12: 'ph:qkv', 13: 'ph:attn', 14: 'ph:outproj', 15: 'ph:LN2'}
with torch.no_grad():
    for dbg in dbgs:
        os.environ['SFB_DBG'] = str(dbg)
        for _ in range(2): ro(x, WL['T_out'])
        torch.cuda.synchronize()
        buf.zero_(); lib.sfb_debug_set_profile(buf.data_ptr(), cap)
        ro(x, WL['T_out']); torch.cuda.synchronize()
        lib.sfb_debug_set_profile(None, 0)
        t = buf.cpu().numpy().astype(np.uint64)
        ev = []
        for role, off in (('P', 2048), ('M', 3072), ('C', 4096)):
            seg = t[off:off + 1024]; seg = seg[seg != 0]
            for v in seg:
                tag = int(v >> np.uint64(48)); clk = int(v & np.uint64(0xFFFFFFFFFFFF))
                ev.append((clk, role, {**COMMON, **TAGS[role]}.get(tag, str(tag))))
        ev.sort()
        t0 = min(c for c, r, n in ev if n == 'LAYER-START')
        print(f'=== dbg={dbg}: {len(ev)} events; cycles since layer start (1 cyc ~ 0.52 ns)')
        for c, r, n in ev:
            print(f'{c - t0:8d} {r} {n}')
