#!/usr/bin/env python
"""Error budget of one rollout step under fp16 operand rounding (CPU emulation, fp64 arithmetic + fp16 rounding at the
points where the tcgen05 engine rounds): which rounding contributes how much to max|out-ref| / max|ref| against
the reference goldens.  Test infrastructure (uses the oracle's LayerNorm); run: python tests/tools/ro_error_budget.py

Result (ro_cfg2, teacher-forced steps 0-5): everything fp16 9.0e-4; weights alone 6.3e-4; LN1 output 3.5e-4; window
slots (in_proj operand) 3.0e-4; out_proj operand 2.8e-4; K 2.7e-4; Q 2.4e-4; attention output 2.3e-4; LN2 output
1.7e-4; V 1.5e-4; FFN hidden 1.4e-4; probabilities 0.6e-4.  in_proj and out_proj see the residual stream at full
magnitude: their two operands are a third of the budget, which is why engine B computes them as three-term
hi/lo products (ro_umma.cu gemm3)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for _p in (ROOT, os.path.join(ROOT, 'tests'), os.path.join(ROOT, 'tests', 'golden')):
    sys.path.insert(0, _p)
import cases
from helpers import golden, rel_max
from oracle import slot_oracle as O
def h16(x): return x.astype(np.float16).astype(np.float64)
def split(x):
    hi = h16(x); return hi + h16(x - hi)
ID = lambda x: x
def step(win, w, pe, K, heads, layers, R):
    g = lambda k: R.get(k, h16)
    W = lambda k: (R['w_io'] if (k in ('in_proj.weight', 'out_proj.weight') and 'w_io' in R) else g('w'))(w[k].astype(np.float64))
    h = g('x')(win) @ W('in_proj.weight').T + w['in_proj.bias'] + pe
    B, L, d = h.shape; dh = d // heads
    for i in range(layers):
        p = f'transformer_encoder.layers.{i}.'
        y = g('ln1')(O.layer_norm(h, w[p+'norm1.weight'].astype(np.float64), w[p+'norm1.bias'].astype(np.float64)))
        qkv = y @ W(p+'self_attn.in_proj_weight').T + w[p+'self_attn.in_proj_bias']
        q, k, v = np.split(qkv, 3, axis=-1)
        sc = 1.4426950408889634/np.sqrt(dh)
        q = g('q')(q*sc)/sc; k = g('k')(k); v = g('v')(v)
        hd = lambda t: t.reshape(B, L, heads, dh).transpose(0,2,1,3)
        q, k, v = hd(q), hd(k), hd(v)
        s = np.einsum('bhid,bhjd->bhij', q, k)/np.sqrt(dh)
        e = np.exp(s - s.max(-1, keepdims=True))
        if R.get('pmode','unnorm') == 'unnorm':
            o = np.einsum('bhij,bhjd->bhid', g('p')(e), v) / g('p')(e).sum(-1, keepdims=True)
        else:
            o = np.einsum('bhij,bhjd->bhid', e, v) / e.sum(-1, keepdims=True)
        o = g('o')(o.transpose(0,2,1,3).reshape(B, L, d))
        h = h + o @ W(p+'self_attn.out_proj.weight').T + w[p+'self_attn.out_proj.bias']
        y = g('ln2')(O.layer_norm(h, w[p+'norm2.weight'].astype(np.float64), w[p+'norm2.bias'].astype(np.float64)))
        f = g('ffh')(np.maximum(y @ W(p+'linear1.weight').T + w[p+'linear1.bias'], 0))
        h = h + f @ W(p+'linear2.weight').T + w[p+'linear2.bias']
    return g('hout')(h[:, -K:]) @ W('out_proj.weight').T + w['out_proj.bias']
L2 = False
for name in ["ro_cfg2", "ro_pack", "ro_cfg5", "ro_cfg3", "ro_cfg4", "ro_tiny", "ro_physion"]:
    c, w, hist = cases.ro_case(name); gd = golden(name); ref = gd['pred_f64']
    B, T_h, K, Ds = hist.shape
    wmax = c['cond_len'] if c['mode'] == 'grow' else T_h
    pe_t = O.sin_pos_enc(wmax, c['d'])
    pe_full = np.repeat(pe_t[0], K, axis=0)[None]
    seq = np.concatenate([hist.astype(np.float64), ref], axis=1)
    def run(R, steps):
        worst = 0
        for s in steps:
            wl = min(T_h + s, wmax)
            win = seq[:, T_h+s-wl:T_h+s].reshape(B, wl*K, Ds)
            out = step(win, w, pe_full[:, -wl*K:], K, c['heads'], c['layers'], R)
            worst = max(worst, (np.linalg.norm(out - ref[:, s]) / np.linalg.norm(ref[:, s])) if L2 else rel_max(out, ref[:, s]))
        return worst
    steps = list(range(min(c['pred_len'], 6)))
    allk = ['x','w','ln1','q','k','v','p','o','ln2','ffh','hout']
    print(name, 'exact', f"{run({k: ID for k in allk}, steps):.2e}", 'all-fp16', f"{run({}, steps):.2e}")
    for k in allk:
        print(f"   only {k:5s} fp16: {run({kk: (h16 if kk == k else ID) for kk in allk}, steps):.2e}   all but {k:5s}: {run({k: ID}, steps):.2e}")
    print('   engine B (x, hout, W_in, W_out as hi+lo):', f"{run({'x': split, 'hout': split, 'w_io': split}, steps):.2e}")
    print('   split x,hout:', f"{run({'x': split, 'hout': split}, steps):.2e}", ' + ln1,ln2 split:', f"{run({'x': split, 'hout': split, 'ln1': split, 'ln2': split}, steps):.2e}",
          ' + w exact:', f"{run({'x': split, 'hout': split, 'ln1': split, 'ln2': split, 'w': ID}, steps):.2e}")
