import sys, numpy as np
sys.path.insert(0,'.'); sys.path.insert(0,'tests/golden')
import cases
from oracle import slot_oracle as O
def f16(x): return x.astype(np.float16).astype(np.float32)
def emul(feats, slots, w, iters, eps=1e-6, tq=f16, wsplit=False):
    w = {k: v.astype(np.float32) for k,v in w.items()}
    B,N,C = feats.shape; D = slots.shape[-1]
    scale = np.float32(D**-0.5)
    xn = O.layer_norm(feats.astype(np.float32), w['norm_inputs.weight'], w['norm_inputs.bias']).astype(np.float32)
    xsum = xn.sum(1); xq = f16(xn)
    Wqk = (scale*np.log2(np.e)) * (w['project_q.1.weight'].T @ w['project_k.weight'])
    Wiv = w['gru.weight_ih'] @ w['project_v.weight']
    def mm(a, b):   # a @ b.T with tail quantisation
        if wsplit:   # activations split hi/lo, weights split hi/lo, 3 products
            ah=f16(a); al=f16(a-ah); bh=f16(b); bl=f16(b-bh)
            return ah@bh.T + al@bh.T + ah@bl.T
        return tq(a) @ tq(b).T
    s = slots.astype(np.float32)
    for it in range(iters):
        qt = mm(O.layer_norm(s, w['project_q.0.weight'], w['project_q.0.bias']).astype(np.float32), Wqk.T.copy())
        hi = f16(qt); lo = f16(qt-hi)
        logits = np.einsum('bnc,bmc->bnm', xq, hi)+np.einsum('bnc,bmc->bnm', xq, lo)
        m = logits.max(-1, keepdims=True); e = np.exp2(logits-m); a = e/e.sum(-1,keepdims=True)
        P = f16(a*1024)
        num = np.einsum('bnm,bnc->bmc', P, xq)/1024 + eps*xsum[:,None,:]
        den = P.sum(1)/1024 + N*eps
        uh = num/den[...,None]
        gi = mm(uh, Wiv) + w['gru.bias_ih']; gh = mm(s, w['gru.weight_hh']) + w['gru.bias_hh']
        r = O._sigmoid(gi[...,:D]+gh[...,:D]); z = O._sigmoid(gi[...,D:2*D]+gh[...,D:2*D]); n = np.tanh(gi[...,2*D:]+r*gh[...,2*D:])
        s = (1-z)*n+z*s
        hid = mm(O.layer_norm(s, w['mlp.0.weight'], w['mlp.0.bias']).astype(np.float32), w['mlp.1.weight']) + w['mlp.1.bias']
        s = s + mm(np.maximum(hid,0), w['mlp.3.weight']) + w['mlp.3.bias']
    return s
for name in ['sa_tiny','sa_cfg1','sa_cfg2','sa_cfg4','sa_ragged']:
    c,w,feats,slots = cases.sa_case(name)
    ref = np.load(f'tests/golden/{name}.npz')['slots_f64']
    for tag,kw in (('tail fp32',dict(tq=lambda x:x)),('tail f16',dict(tq=f16)),('tail split3',dict(wsplit=True))):
        s = emul(feats, slots, w, c['iters'], **kw)
        print(f'{name:10s} {tag:12s} max-rel {np.abs(s-ref).max()/np.abs(ref).max():.2e}')
print('--- 2-term variants')
def emul2(feats, slots, w, iters, mode):
    def f(a,b):
        ah=f16(a); al=f16(a-ah); bh=f16(b); bl=f16(b-bh)
        if mode=='A2': return ah@bh.T + al@bh.T
        if mode=='W2': return ah@bh.T + ah@bl.T
    import types
    g = emul.__globals__
    return None
import types
src = open(__file__).read()
