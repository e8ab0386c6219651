"""Golden-case table + portable deterministic weight / input generators.

Everything is generated with ``numpy.random.RandomState`` (bit-stable across
numpy versions and machines), so a golden file only has to store the
*reference outputs*; weights and inputs are regenerated from the seed on the
GPU box, where /root/reference does not exist.

Used by: tests/golden/make_golden.py (runs the real reference, build container
only), tests/*, bench.py and __graft_entry__.smoke().
"""
import numpy as np

# --------------------------------------------------------------------------- #
# case tables (shapes follow BASELINE.json configs)
# --------------------------------------------------------------------------- #
# Slot Attention: B frames, N pixels, C in-features, D slot size, Dm mlp hidden,
# K slots, iters; mask=True -> SlotAttentionWMask (steve.py) variant.
SA_CASES = {
    'sa_tiny':  dict(B=2, N=256, C=128, D=128, Dm=256, K=4, iters=3, mask=False, seed=11),
    'sa_cfg1':  dict(B=4, N=4096, C=128, D=128, Dm=256, K=5, iters=3, mask=False, seed=12),
    'sa_cfg2':  dict(B=8, N=4096, C=128, D=128, Dm=256, K=6, iters=2, mask=False, seed=13),
    'sa_cfg3':  dict(B=4, N=4096, C=128, D=128, Dm=256, K=7, iters=2, mask=False, seed=14),
    'sa_cfg4':  dict(B=2, N=4096, C=192, D=192, Dm=384, K=8, iters=2, mask=True, seed=15),
    'sa_ragged': dict(B=3, N=1000, C=128, D=128, Dm=256, K=6, iters=2, mask=True, seed=16),
    'sa_one':   dict(B=1, N=64, C=128, D=128, Dm=256, K=1, iters=1, mask=True, seed=17),
}

# Rollout: B clips, T_h history frames, K slots, Ds slot size, d model width,
# F ffn width, layers, heads, pred_len; mode slide|grow (grow => cond_len).
RO_CASES = {
    'ro_tiny':    dict(B=2, T_h=2, K=3, Ds=128, d=128, F=512, layers=2, heads=8,
                       pred_len=3, mode='slide', cond_len=None, seed=21),
    'ro_cfg2':    dict(B=4, T_h=6, K=6, Ds=128, d=128, F=512, layers=4, heads=8,
                       pred_len=10, mode='slide', cond_len=None, seed=22),
    'ro_cfg3':    dict(B=2, T_h=6, K=7, Ds=128, d=256, F=1024, layers=4, heads=8,
                       pred_len=44, mode='slide', cond_len=None, seed=23),
    'ro_cfg4':    dict(B=2, T_h=4, K=8, Ds=192, d=256, F=1024, layers=8, heads=8,
                       pred_len=20, mode='slide', cond_len=None, seed=24),
    'ro_cfg5':    dict(B=2, T_h=1, K=6, Ds=128, d=256, F=1024, layers=8, heads=8,
                       pred_len=64, mode='grow', cond_len=6, seed=25),
    'ro_physion': dict(B=1, T_h=15, K=6, Ds=192, d=256, F=1024, layers=8, heads=8,
                       pred_len=10, mode='slide', cond_len=None, seed=26),
    'ro_pack':    dict(B=7, T_h=6, K=6, Ds=128, d=128, F=512, layers=4, heads=8,
                       pred_len=4, mode='slide', cond_len=None, seed=27),
}


# --------------------------------------------------------------------------- #
# generators
# --------------------------------------------------------------------------- #
def _lin(rs, out_f, in_f, gain=1.0):
    """Weight ~ U(-a, a), a = gain*sqrt(3/in_f): unit-variance-preserving."""
    a = gain * np.sqrt(3.0 / in_f)
    return rs.uniform(-a, a, size=(out_f, in_f)).astype(np.float32)


def _vec(rs, n, scale=0.1, base=0.0):
    return (base + scale * rs.standard_normal(n)).astype(np.float32)


def make_sa_weights(C, D, Dm, seed):
    """SlotAttention state_dict (reference savi.py:19-54 parameter set)."""
    rs = np.random.RandomState(seed)
    w = {}
    w['norm_inputs.weight'] = _vec(rs, C, 0.1, 1.0)
    w['norm_inputs.bias'] = _vec(rs, C, 0.1)
    w['project_q.0.weight'] = _vec(rs, D, 0.1, 1.0)
    w['project_q.0.bias'] = _vec(rs, D, 0.1)
    w['project_q.1.weight'] = _lin(rs, D, D, 2.0)
    w['project_k.weight'] = _lin(rs, D, C, 2.0)
    w['project_v.weight'] = _lin(rs, D, C)
    w['gru.weight_ih'] = _lin(rs, 3 * D, D)
    w['gru.weight_hh'] = _lin(rs, 3 * D, D)
    w['gru.bias_ih'] = _vec(rs, 3 * D)
    w['gru.bias_hh'] = _vec(rs, 3 * D)
    w['mlp.0.weight'] = _vec(rs, D, 0.1, 1.0)
    w['mlp.0.bias'] = _vec(rs, D, 0.1)
    w['mlp.1.weight'] = _lin(rs, Dm, D)
    w['mlp.1.bias'] = _vec(rs, Dm)
    w['mlp.3.weight'] = _lin(rs, D, Dm)
    w['mlp.3.bias'] = _vec(rs, D)
    return w


def make_sa_inputs(B, N, C, D, K, seed, spatial=True):
    """feats [B,N,C] with per-pixel offsets/scales (exercises LN), slots [B,K,D]."""
    rs = np.random.RandomState(seed + 1000)
    feats = rs.standard_normal((B, N, C)).astype(np.float32)
    if spatial:
        feats *= (0.5 + rs.uniform(0, 1.5, size=(B, N, 1))).astype(np.float32)
        feats += (0.7 * rs.standard_normal((B, N, 1))).astype(np.float32)
    slots = rs.standard_normal((B, K, D)).astype(np.float32)
    return feats, slots


def make_ro_weights(Ds, d, F, layers, seed):
    """SlotRollouter state_dict minus enc_t_pe (reference slotformer.py:51-83)."""
    rs = np.random.RandomState(seed)
    w = {}
    w['in_proj.weight'] = _lin(rs, d, Ds)
    w['in_proj.bias'] = _vec(rs, d)
    for i in range(layers):
        p = f'transformer_encoder.layers.{i}.'
        w[p + 'self_attn.in_proj_weight'] = _lin(rs, 3 * d, d, 1.5)
        w[p + 'self_attn.in_proj_bias'] = _vec(rs, 3 * d)
        w[p + 'self_attn.out_proj.weight'] = _lin(rs, d, d, 0.7)
        w[p + 'self_attn.out_proj.bias'] = _vec(rs, d)
        w[p + 'linear1.weight'] = _lin(rs, F, d)
        w[p + 'linear1.bias'] = _vec(rs, F)
        w[p + 'linear2.weight'] = _lin(rs, d, F, 0.7)
        w[p + 'linear2.bias'] = _vec(rs, d)
        w[p + 'norm1.weight'] = _vec(rs, d, 0.1, 1.0)
        w[p + 'norm1.bias'] = _vec(rs, d, 0.1)
        w[p + 'norm2.weight'] = _vec(rs, d, 0.1, 1.0)
        w[p + 'norm2.bias'] = _vec(rs, d, 0.1)
    w['out_proj.weight'] = _lin(rs, Ds, d, 0.5)
    w['out_proj.bias'] = _vec(rs, Ds)
    return w


def make_ro_inputs(B, T_h, K, Ds, seed):
    rs = np.random.RandomState(seed + 1000)
    return rs.standard_normal((B, T_h, K, Ds)).astype(np.float32)


def sa_case(name):
    c = dict(SA_CASES[name])
    w = make_sa_weights(c['C'], c['D'], c['Dm'], c['seed'])
    feats, slots = make_sa_inputs(c['B'], c['N'], c['C'], c['D'], c['K'], c['seed'])
    return c, w, feats, slots


def ro_case(name):
    c = dict(RO_CASES[name])
    w = make_ro_weights(c['Ds'], c['d'], c['F'], c['layers'], c['seed'])
    hist = make_ro_inputs(c['B'], c['T_h'], c['K'], c['Ds'], c['seed'])
    return c, w, hist
