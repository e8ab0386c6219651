import sys, numpy as np
sys.path.insert(0,'.'); sys.path.insert(0,'tests/golden')
import cases
from oracle import slot_oracle as O
def f16(x): return x.astype(np.float16).astype(x.dtype)
for name in ['ro_tiny','ro_cfg2','ro_cfg3','ro_cfg5']:
    c,w,hist = cases.ro_case(name)
    g = np.load(f'tests/golden/{name}.npz'); w=dict(w); w['enc_t_pe']=g['enc_t_pe']
    ref = g['pred_f64']
    for tag,fn in (('bf16',O.to_bf16),('f16',f16),('tf32',O.to_tf32)):
        out = O.rollout(hist,w,c['pred_len'],c['heads'],c['layers'],mode=c['mode'],cond_len=c['cond_len'],operand_round=fn)
        e = np.abs(out-ref)
        per_step = e.reshape(e.shape[0],e.shape[1],-1).max(-1).max(0)/np.abs(ref).max()
        print(name,tag,'max-rel %.2e'%(e.max()/np.abs(ref).max()),'step0 %.2e last %.2e'%(per_step[0],per_step[-1]))
