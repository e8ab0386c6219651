#!/usr/bin/env python
"""Golden for the offline rollout driver: the UNMODIFIED reference
``rollout_video_slots`` (slotformer/video_prediction/rollout_clevrer_slots.py:19-66) run with a deterministic
stand-in model, which pins the frame-offset index arithmetic (which observed frames feed which offset, how the
per-offset predictions interleave, zero padding to 160 frames).  Build-container only (needs /root/reference;
``Tensor.cuda`` / ``torch.cuda.device_count`` are patched because this container has no GPU)."""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, HERE)
import make_golden  # noqa: E402

K, D = 3, 4
CASES = {'off1': dict(history_len=6, frame_offset=1), 'off2': dict(history_len=6, frame_offset=2),
         'off3': dict(history_len=6, frame_offset=3)}
NAMES = ['video_%02d' % i for i in range(5)]


def fake_rollout(hist, pred_len):
    """Deterministic stand-in for SlotRollouter.forward: step s of a clip = mean of its history window + (s + 1) / 8
    (depends on every history frame and on the step, so any index slip changes the result)."""
    base = hist.mean(dim=1)
    return torch.stack([base + (s + 1) / 8.0 for s in range(pred_len)], dim=1)


def make_pre_slots(seed=3):
    rs = np.random.RandomState(seed)
    return {n: rs.standard_normal((128, K, D)).astype(np.float32) for n in NAMES}


class FakeSlotFormer:
    """What rollout_video_slots touches: .eval(), .module.rollout_len, __call__({'slots': x})['pred_slots']."""

    def __init__(self, history_len):
        self.history_len = history_len
        self.rollout_len = None
        self.module = self

    def eval(self):
        return self

    def __call__(self, data_dict):
        x = data_dict['slots']
        return {'pred_slots': fake_rollout(x[:, :self.history_len], self.rollout_len)}


def main():
    make_golden.import_reference()
    sys.path.insert(0, os.path.join(make_golden.REF, 'slotformer', 'video_prediction'))
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.cuda.device_count = lambda: 2
    import rollout_clevrer_slots as R
    out = {}
    pre = make_pre_slots()
    for name, c in CASES.items():
        R.params = types.SimpleNamespace(input_frames=c['history_len'], frame_offset=c['frame_offset'])
        res = R.rollout_video_slots(FakeSlotFormer(c['history_len']), pre)
        out[name] = np.stack([res[n] for n in NAMES])
    path = os.path.join(HERE, 'offline.npz')
    np.savez_compressed(path, **out)
    print('wrote', path, os.path.getsize(path) // 1024, 'KiB', {k: v.shape for k, v in out.items()})


if __name__ == '__main__':
    main()
