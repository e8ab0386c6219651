"""Offline drivers: frame-offset rollout index arithmetic against the unmodified reference
function (tests/golden/offline.npz), file formats, sharding; on the GPU the driver with the real rollout kernel
against the numpy oracle."""
import os
import sys

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, 'golden'))
import make_golden_offline as G  # noqa: E402
from slotformer_b200 import offline  # noqa: E402

GOLD = np.load(os.path.join(HERE, 'golden', 'offline.npz'))


@pytest.mark.parametrize('name', list(G.CASES))
def test_frame_offset_rollout_matches_reference_driver(name):
    c = G.CASES[name]
    pre = G.make_pre_slots()
    for bv in (2, 64):          # batching must not matter
        res = offline.rollout_video_slots(G.fake_rollout, pre, c['history_len'], c['frame_offset'], obs_frames=128,
                                          target_len=160, batch_videos=bv, device='cpu')
        got = np.stack([res[n] for n in G.NAMES])
        assert got.dtype == np.float32 and got.shape == GOLD[name].shape
        assert np.array_equal(got, GOLD[name])          # index arithmetic: bit exact


def test_uneven_offsets_and_errors():
    # 7 future frames at offset 3: sub-sequences need 3, 2, 2 steps
    pre = {'v': np.arange(20 * 2 * 2, dtype=np.float32).reshape(20, 2, 2)}
    res = offline.rollout_video_slots(G.fake_rollout, pre, 2, 3, obs_frames=20, target_len=27, device='cpu')['v']
    for i in range(7):
        o, s = i % 3, i // 3
        start = 20 - 2 * 3 + o
        hist = pre['v'][start::3][:2]
        assert np.allclose(res[20 + i], hist.mean(0) + (s + 1) / 8.0)
    with pytest.raises(ValueError):
        offline.offset_starts(obs_frames=4, history_len=6, frame_offset=1)


def test_file_formats_and_sharding(tmp_path):
    rs = np.random.RandomState(0)
    table = {'train': {f'{i}.mp4': rs.standard_normal((7, 3, 4)).astype(np.float32) for i in range(5)},
             'val': {'a.mp4': rs.standard_normal((7, 3, 4)).astype(np.float32)}}
    path = str(tmp_path / 'sub' / 'slots.pkl')
    offline.dump_slots(table, path)
    import pickle
    with open(path, 'rb') as f:                       # plain pickle of {split: {name: float32[T,K,D]}}
        back = pickle.load(f)
    assert set(back) == {'train', 'val'} and all(np.array_equal(back['train'][k], v) for k, v in table['train'].items())
    assert offline.load_slots(path)['val']['a.mp4'].dtype == np.float32
    with pytest.raises(ValueError):
        offline.dump_slots({'train': {'x': np.zeros((2, 2), np.float32)}}, path)
    offline.save_phyre_sample(str(tmp_path / 'phyre'), 12, np.ones((9, 6, 8), np.float32), vid_len=5)
    assert np.load(str(tmp_path / 'phyre' / '000012.npy')).shape == (5, 6, 8)
    names = [f'v{i}' for i in range(11)]
    parts = [offline.shard_names(names, r, 4) for r in range(4)]
    assert sum(parts, []) == names and max(map(len, parts)) - min(map(len, parts)) <= 1
    merged = offline.merge_shards([{n: np.zeros(1) for n in p} for p in parts])
    assert list(merged) == names
    with pytest.raises(ValueError):
        offline.merge_shards([{'a': 1}, {'a': 2}])


def test_extract_video_slots_batches_equal_lengths():
    calls = []

    class M(torch.nn.Module):
        def forward(self, d):
            calls.append(tuple(d['img'].shape))
            v = d['img']
            return {'post_slots': v.mean(dim=(2, 3, 4))[:, :, None, None].expand(-1, -1, 2, 3)}

    vids = {f'v{i}': torch.full((4 if i < 3 else 6, 3, 2, 2), float(i)) for i in range(5)}
    out = offline.extract_video_slots(M(), vids.__getitem__, list(vids), batch_videos=2, device='cpu')
    assert calls == [(2, 4, 3, 2, 2), (1, 4, 3, 2, 2), (2, 6, 3, 2, 2)]
    assert out['v4'].shape == (6, 2, 3) and out['v4'].dtype == np.float32 and np.all(out['v4'] == 4.0)


@pytest.mark.gpu
def test_offline_rollout_with_kernel_matches_oracle():
    """CLEVRER-style job at small scale through the tcgen05 rollout kernel: 8 videos, offset 2, 6 -> 10 frames."""
    import cases
    from helpers import golden, ro_module, rel_max
    from oracle import slot_oracle as O
    c, w, _ = cases.ro_case('ro_cfg2')
    g = golden('ro_cfg2')
    m = ro_module(c, w, 'cuda:0', enc_t_pe=g['enc_t_pe'])
    rs = np.random.RandomState(5)
    pre = {f'v{i}': rs.standard_normal((24, c['K'], c['Ds'])).astype(np.float32) for i in range(8)}
    res = offline.rollout_video_slots(m, pre, c['T_h'], 2, obs_frames=24, target_len=34, batch_videos=8)
    w2 = dict(w)
    w2['enc_t_pe'] = g['enc_t_pe']
    for n in ('v0', 'v7'):
        assert np.array_equal(res[n][:24], pre[n])
        for o in range(2):
            hist = pre[n][24 - 12 + o::2][:6][None]
            ref = O.rollout(hist, w2, 5, c['heads'], c['layers'])[0]
            assert rel_max(res[n][24 + o::2], ref) < 4e-3


def test_phyre_split_ranges_and_resume(tmp_path):
    """extract_phyre_slots.py:41-53 arithmetic: equal shares, the last job takes the remainder; a restarted job
    re-does the last file it wrote and skips the rest."""
    total, n = 103, 8
    ranges = [offline.phyre_split_range(total, s, n) for s in range(n)]
    assert ranges[0] == (0, 12) and ranges[-1] == (84, 103)
    assert all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))          # contiguous, disjoint, complete
    assert offline.phyre_split_range(total, -1, n) == (0, total)
    with pytest.raises(ValueError):
        offline.phyre_split_range(total, 8, 8)

    class FakeSavi(torch.nn.Module):
        testing = True

        def forward(self, d):               # [B, T, 3, H, W] -> post_slots [B, T, 2, 4] = per-frame pixel mean + slot id
            m = d['img'].mean(dim=(2, 3, 4))
            return {'post_slots': m[:, :, None, None] + torch.arange(2.)[None, None, :, None] + torch.zeros(4)}

    def get_sample(i):
        return np.full((5, 3, 4, 4), float(i), dtype=np.float32), 3 + i % 3

    root = str(tmp_path / 'slots')
    calls = []

    def counting(i):
        calls.append(i)
        return get_sample(i)

    wrote = offline.extract_phyre_job(FakeSavi(), counting, 20, root, split=1, total_split=3, batch_size=4, device='cpu')
    assert wrote == list(range(6, 12))
    for i in wrote:
        s = np.load(os.path.join(root, f'{i:06d}.npy'))
        assert s.dtype == np.float32 and s.shape == (3 + i % 3, 2, 4) and np.allclose(s[:, 1], i + 1)
    os.remove(os.path.join(root, '000010.npy'))                # simulate a job killed after sample 9
    os.remove(os.path.join(root, '000011.npy'))
    calls.clear()
    wrote = offline.extract_phyre_job(FakeSavi(), counting, 20, root, split=1, total_split=3, batch_size=4, device='cpu')
    assert wrote == [9, 10, 11] and calls == [9, 10, 11]
