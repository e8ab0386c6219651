#!/usr/bin/env python
"""Generate golden vectors by running the UNMODIFIED reference modules.

Build-container only: imports /root/reference (read-only) through the ``nerv``
shim plus empty stand-ins for data-side deps the model files never call
(pycocotools, phyre, ...).  Weights and inputs come from tests/golden/cases.py
(seeded numpy), are loaded into the reference ``SlotAttention`` /
``SlotAttentionWMask`` / ``SlotRollouter`` / ``SingleStepSlotRollouter`` via
``load_state_dict`` and run in eval + no_grad, fp32 and fp64 CPU.  Only the
outputs are stored (tests/golden/<case>.npz).

    python tests/golden/make_golden.py            # all cases
    python tests/golden/make_golden.py sa_tiny    # some cases

The produced .npz files are committed; the GPU box never runs this script.
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get('SLOTFORMER_REFERENCE', '/root/reference')
sys.path.insert(0, REPO)
sys.path.insert(0, HERE)

from slotformer_b200.compat import install_nerv_shim  # noqa: E402
import cases  # noqa: E402


class _Anything(types.ModuleType):
    """Module stand-in: any attribute is another stand-in / dummy callable."""

    def __getattr__(self, name):
        if name.startswith('__'):
            raise AttributeError(name)
        child = _Anything(f'{self.__name__}.{name}')
        setattr(self, name, child)
        return child

    def __call__(self, *a, **k):
        return None


def import_reference():
    install_nerv_shim()
    import nerv.utils as nu
    for name in ('glob_all', 'read_all_lines', 'read_img', 'VideoReader',
                 'AverageMeter', 'MeanMetric', 'save_video', 'batch_gather',
                 'batch_cat_vec', 'load_obj', 'dump_obj'):
        if not hasattr(nu, name):
            setattr(nu, name, lambda *a, **k: None)
    import nerv.training as nt
    for name in ('BaseMethod', 'BaseDataModule', 'CosineAnnealingWarmupRestarts'):
        if not hasattr(nt, name):
            setattr(nt, name, type(name, (), {}))
    for missing in ('pycocotools', 'pycocotools.mask', 'phyre', 'lpips',
                    'skimage', 'skimage.metrics', 'moviepy', 'moviepy.editor',
                    'wandb'):
        try:
            __import__(missing)
        except Exception:
            sys.modules[missing] = _Anything(missing)
    sys.path.append(REF)
    from slotformer.base_slots.models.savi import SlotAttention
    from slotformer.base_slots.models.steve import SlotAttentionWMask
    from slotformer.video_prediction.models.slotformer import SlotRollouter
    from slotformer.video_prediction.models.single_step_slotformer import \
        SingleStepSlotRollouter
    return SlotAttention, SlotAttentionWMask, SlotRollouter, SingleStepSlotRollouter


def _load(module, weights, dtype, extra_ok=()):
    sd = {k: torch.from_numpy(np.asarray(v)).to(dtype) for k, v in weights.items()}
    missing, unexpected = module.load_state_dict(sd, strict=False)
    assert not unexpected, unexpected
    assert set(missing) <= set(extra_ok), missing
    return module.to(dtype).eval()


def gen_sa(name, SA, SAM):
    c, w, feats, slots = cases.sa_case(name)
    cls = SAM if c['mask'] else SA
    out = {}
    for tag, dt in (('f32', torch.float32), ('f64', torch.float64)):
        m = cls(in_features=c['C'], num_iterations=c['iters'], num_slots=c['K'],
                slot_size=c['D'], mlp_hidden_size=c['Dm'])
        m = _load(m, w, dt)
        with torch.no_grad():
            r = m(torch.from_numpy(feats).to(dt), torch.from_numpy(slots).to(dt))
        if c['mask']:
            out[f'slots_{tag}'] = r[0].numpy()
            out[f'mask_{tag}'] = r[1].numpy()
        else:
            out[f'slots_{tag}'] = r.numpy()
    # provenance: state_dict key order/shapes of the reference module
    out['keys'] = np.array(list(m.state_dict().keys()))
    return out


def gen_ro(name, RO, SSRO):
    c, w, hist = cases.ro_case(name)
    kw = dict(num_slots=c['K'], slot_size=c['Ds'], history_len=c['T_h'],
              d_model=c['d'], num_layers=c['layers'], num_heads=c['heads'],
              ffn_dim=c['F'])
    out = {}
    for tag, dt in (('f32', torch.float32), ('f64', torch.float64)):
        if c['mode'] == 'grow':
            m = SSRO(cond_len=c['cond_len'], **kw)
        else:
            m = RO(**kw)
        pe_ref = m.enc_t_pe.detach().clone()
        m = _load(m, w, dt, extra_ok=('enc_t_pe',))
        with torch.no_grad():
            m.enc_t_pe.copy_(pe_ref.to(dt))
            r = m(torch.from_numpy(hist).to(dt), c['pred_len'])
        out[f'pred_{tag}'] = r.numpy()
    out['enc_t_pe'] = pe_ref.numpy()
    out['keys'] = np.array(list(m.state_dict().keys()))
    # which torch path produced it (eval + no_grad => fused encoder fast path)
    out['torch_version'] = np.array(torch.__version__)
    return out


def main(argv):
    torch.manual_seed(0)
    torch.set_num_threads(8)
    SA, SAM, RO, SSRO = import_reference()
    want = argv or (list(cases.SA_CASES) + list(cases.RO_CASES))
    for name in want:
        out = gen_sa(name, SA, SAM) if name in cases.SA_CASES else gen_ro(name, RO, SSRO)
        path = os.path.join(HERE, f'{name}.npz')
        np.savez_compressed(path, **out)
        sizes = {k: v.shape for k, v in out.items() if v.dtype.kind == 'f'}
        print(f'{name}: wrote {os.path.getsize(path)/1024:.1f} KiB {sizes}')


if __name__ == '__main__':
    main(sys.argv[1:])
