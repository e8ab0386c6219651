#!/usr/bin/env python
"""Goldens for the SAVi slot transition: the UNMODIFIED reference StoSAVi's
predictor / kernel_dist_layer / _sample_dist chain (savi.py:393-410, predictor.py:20-113) on CPU (evaluated in fp64, stored as fp32, with the error of the reference's own fp32 evaluation beside it),
with torch.randn_like replaced by the case's fixed noise.  Build-container only (needs /root/reference)."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, HERE)
import make_golden  # noqa: E402  (reference import machinery)
import transition_cases as TC  # noqa: E402


def run(ref, name, dtype):
    kw, B, steps, _ = TC.CASES[name]
    prev, noise = TC.inputs(name)
    ref = ref.to(dtype)
    dists, inits = [], []
    it = iter(torch.from_numpy(noise).to(dtype))
    orig = torch.randn_like
    torch.randn_like = lambda t, **k: next(it).to(t.dtype)      # _sample_dist draws eps = randn_like(mu) (savi.py:362)
    try:
        with torch.no_grad():
            if hasattr(ref.predictor, 'reset'):
                ref.predictor.reset()
            for t in range(steps + 1):
                # exactly the body of the frame loop, savi.py:394-403
                if t == 0:
                    latents = ref.init_latents.repeat(B, 1, 1)
                else:
                    latents = ref.predictor(torch.from_numpy(prev[t - 1]).to(dtype))
                dist = ref.kernel_dist_layer(latents)
                kernels = ref._sample_dist(dist)
                if ref.kld_method == 'none':
                    next(it)                                     # keep the noise stream aligned per frame
                dists.append(dist.double().numpy())
                inits.append(kernels.double().numpy())
    finally:
        torch.randn_like = orig
    out = {'dist': np.stack(dists), 'init': np.stack(inits)}
    hs = getattr(ref.predictor, 'hidden_state', None)
    if hs is not None:
        out['h'] = hs[0].double().numpy()
        out['c'] = hs[1].double().numpy()
    return out


def main():
    make_golden.import_reference()
    from slotformer.base_slots.models import StoSAVi as RefSAVi
    out = {}
    for name in TC.CASES:
        r32 = run(TC.build(RefSAVi, name), name, torch.float32)
        r64 = run(TC.build(RefSAVi, name), name, torch.float64)
        for k, v in r64.items():
            out[f'{name}.{k}'] = v.astype(np.float32)           # fp64 evaluation of the reference, stored as fp32
            # how far the reference's own fp32 evaluation is from it (the scale of the test tolerance)
            out[f'{name}.{k}_f32err'] = np.float64(np.abs(r32[k] - v).max() / np.abs(v).max())
    path = os.path.join(HERE, 'transition.npz')
    np.savez_compressed(path, **out)
    print('wrote', path, os.path.getsize(path) // 1024, 'KiB')


if __name__ == '__main__':
    main()
