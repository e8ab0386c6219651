import sys, os, numpy as np
sys.path.insert(0,'.'); sys.path.insert(0,'tests/golden')
import cases
from oracle import slot_oracle as O
def f16(x): return x.astype(np.float16).astype(np.float32)
def bf16(x): return O.to_bf16(x.astype(np.float32))
def emul(feats, slots, w, iters, eps=1e-6, q=f16, qq=f16, split_q=False):
    w = {k: v.astype(np.float32) for k,v in w.items()}
    B,N,C = feats.shape; D = slots.shape[-1]
    scale = np.float32(D**-0.5)
    x = feats.astype(np.float32)
    xn = O.layer_norm(x, w['norm_inputs.weight'], w['norm_inputs.bias']).astype(np.float32)
    xsum = xn.sum(1)   # [B,C]
    xq = q(xn)
    Wqk = (scale*np.log2(np.e)) * (w['project_q.1.weight'].T @ w['project_k.weight'])   # [D,C]
    Wiv = w['gru.weight_ih'] @ w['project_v.weight']   # [3D, C]
    s = slots.astype(np.float32)
    mask=None
    for it in range(iters):
        qt = O.layer_norm(s, w['project_q.0.weight'], w['project_q.0.bias']).astype(np.float32) @ Wqk  # [B,K,C]
        if split_q:
            hi = qq(qt); lo = qq(qt-hi); logits = np.einsum('bnc,bmc->bnm', xq, hi)+np.einsum('bnc,bmc->bnm', xq, lo)
        else:
            logits = np.einsum('bnc,bmc->bnm', xq, qq(qt))
        m = logits.max(-1, keepdims=True)
        e = np.exp2(logits-m)
        a = e/e.sum(-1,keepdims=True)
        mask = a.transpose(0,2,1)
        P = q(a*1024)
        num = np.einsum('bnm,bnc->bmc', P, xq)/1024 + eps*xsum[:,None,:]
        den = P.sum(1)/1024 + N*eps
        uh = num/den[...,None]
        gi = uh @ Wiv.T + w['gru.bias_ih']; gh = s @ w['gru.weight_hh'].T + w['gru.bias_hh']
        r = O._sigmoid(gi[...,:D]+gh[...,:D]); z = O._sigmoid(gi[...,D:2*D]+gh[...,D:2*D]); n = np.tanh(gi[...,2*D:]+r*gh[...,2*D:])
        s = (1-z)*n+z*s
        hid = O.layer_norm(s, w['mlp.0.weight'], w['mlp.0.bias']) @ w['mlp.1.weight'].T + w['mlp.1.bias']
        s = s + np.maximum(hid,0) @ w['mlp.3.weight'].T + w['mlp.3.bias']
    return s, mask
for name in cases.SA_CASES:
    c,w,feats,slots = cases.sa_case(name)
    g = np.load(f'tests/golden/{name}.npz')
    ref = g['slots_f64']
    for tag,q,qq,sp in (('fp32',lambda x:x,lambda x:x,False),('f16',f16,f16,False),('f16+splitq',f16,f16,True),('bf16',bf16,bf16,False)):
        s,mask = emul(feats, slots, w, c['iters'], q=q, qq=qq, split_q=sp)
        err = np.abs(s-ref); rel = err.max()/np.abs(ref).max()
        erel = (err/(np.abs(ref)+1e-3*np.abs(ref).max())).max()
        ms = ''
        if c['mask']: ms = ' mask maxabs %.2e'%np.abs(mask-g['mask_f64']).max()
        print(f'{name:10s} {tag:11s} max-rel {rel:.2e} elem-rel {erel:.2e}{ms}')
