#!/usr/bin/env python
"""Workload for compute-sanitizer: Slot Attention (tcgen05 and mma.sync passes, ragged + mask cases), encoder tail,
rollout goldens incl. ro_cfg5 at B = 256 (two waves of CTAs), decoder epilogue -- every kernel family once."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests')); sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden'))
import numpy as np, torch
import cases
from helpers import golden, rel_max, ro_module, sa_module
from slotformer_b200 import engine
dev = 'cuda:0'
big = int(os.environ.get('RO_B', '256'))
ONLY = os.environ.get('SAN_ONLY', '')
with torch.no_grad():
    for name in (() if ONLY else ('sa_cfg2', 'sa_ragged', 'sa_cfg4', 'sa_one')):
        c, w, feats, slots = cases.sa_case(name)
        for flags in (0, engine.SFB_SA_NO_TCGEN05):
            m = sa_module(c, w, dev); m.engine_flags = flags
            out = m(torch.from_numpy(feats).to(dev), torch.from_numpy(slots).to(dev))
            out = out[0] if isinstance(out, tuple) else out
            print(name, flags, f'{rel_max(out.cpu().numpy(), golden(name)["slots_f64"]):.2e}', flush=True)
    for name in (() if ONLY else ('ro_cfg2', 'ro_cfg5', 'ro_physion')):
        c, w, hist = cases.ro_case(name)
        g = golden(name)
        m = ro_module(c, w, dev, enc_t_pe=g['enc_t_pe'])
        x = torch.from_numpy(hist).to(dev)
        if name == 'ro_cfg5':
            x = torch.cat([x, torch.randn((big - x.shape[0],) + x.shape[1:], device=dev)], 0)
        out = m(x, c['pred_len'])[:hist.shape[0]]
        print(name, f'{rel_max(out.cpu().numpy(), g["pred_f64"]):.2e}', flush=True)
    import wrapper_cases as W
    from slotformer_b200.base_slots.models import StoSAVi
    if not ONLY:
        x = torch.randn((6, 6, 4, 64, 64), device=dev)
        engine.decode_combine(x, want_seg=True)
        tail = engine.EncoderTailEngine()
        sv = W.build_savi(StoSAVi).to(dev).eval()
        named = dict(sv.named_parameters())
        t = tail.forward(torch.randn((3, 64, 20, 20), device=dev), {k: named[k] for k in engine.ENC_TAIL_KEYS}, 128)
        print('tail', bool(torch.isfinite(t.data.float()).all()), flush=True)
    # slot transition kernel (clusters of 8 / 4 / 2 / 1 CTAs per clip, DSMEM writes)
    if os.environ.get('SAN_ONLY', '') in ('', 'transition'):
        import transition_cases as TC
        for name, tile in (('tr_obj3d', 1), ('tr_obj3d', 6), ('tr_postln', 10), ('tr_clevrer', 40), ('tr_plain', 1)):
            kw, B0, steps, _ = TC.CASES[name]
            tm = TC.build(StoSAVi, name).to(dev)
            spec = tm._transition_spec()
            eng = engine.TransitionEngine()
            prev, noise = TC.inputs(name)
            B = B0 * tile
            state = None
            for t in range(steps + 1):
                nz = torch.from_numpy(np.concatenate([noise[t]] * tile)).to(dev) if tm.kld_method != 'none' else None
                if t == 0:
                    dist, init, _ = eng.forward(spec, tm.init_latents.detach(), False, B, None, nz)
                else:
                    dist, init, state = eng.forward(spec, torch.from_numpy(np.concatenate([prev[t - 1]] * tile)).to(dev), True, B, state, nz)
            g = np.load(os.path.join(ROOT, 'tests', 'golden', 'transition.npz'))
            print(name, tile, f'{rel_max(dist.cpu().numpy(), np.concatenate([g[name + ".dist"][steps]] * tile)):.2e}', flush=True)
torch.cuda.synchronize()
print('done')
