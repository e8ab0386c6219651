"""Driver-facing contracts that can be checked without a GPU: the reference arm of bench.py prints exactly one JSON
line with the agreed keys, and the sharded offline job (one process per rank, files as the only rendezvous) produces
the same table as a single process."""
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_bench_reference_arm_prints_one_json_line():
    env = dict(os.environ, OMP_NUM_THREADS='1')          # as torchrun exports it: the arm must still use every core
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--steps', '1',
                          '--warmup', '1'], capture_output=True, text=True, env=env, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d['impl'] == 'reference' and d['metric'] == 'video_clip_frames_per_sec' and d['unit'] == 'frames/s'
    assert d['higher_is_better'] is True and d['value'] > 0 and d['gpu_launches'] == 0
    # the arm runs the UNMODIFIED reference modules (oracle/_ref), full-size steps, nothing extrapolated
    assert d['cpu_baseline']['kind'] == 'reference' and d['cpu_baseline']['cores'] == os.cpu_count()
    assert 'unmodified reference' in d['cpu_baseline']['sample'] and 'extrapolat' not in json.dumps(d)
    assert d['e2e'] == {'value': d['value'], 'unit': 'frames/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}
    assert d['config']['workload'].startswith('OBJ3D SlotFormer rollout, B=64')


_WORKER = r'''
import os, sys, numpy as np, torch
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, os.path.join(sys.argv[1], 'tests', 'golden'))
import make_golden_offline as G
from slotformer_b200 import offline
rank, world, out = int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]
pre = G.make_pre_slots()
names = offline.shard_names(list(pre), rank, world)
res = offline.rollout_video_slots(G.fake_rollout, {n: pre[n] for n in names}, 6, 2, obs_frames=128, target_len=160,
                                  batch_videos=2, device='cpu')
offline.dump_slots({'val': res}, out + '.rank%d' % rank)
'''


def test_sharded_offline_job_equals_single_process(tmp_path):
    sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden'))
    import make_golden_offline as G
    from slotformer_b200 import offline
    out = str(tmp_path / 'rollout_slots.pkl')
    world = 2
    procs = [subprocess.Popen([sys.executable, '-c', _WORKER, ROOT, str(r), str(world), out]) for r in range(world)]
    assert all(p.wait(timeout=300) == 0 for p in procs)
    parts = [offline.load_slots(out + '.rank%d' % r)['val'] for r in range(world)]
    assert sorted(map(len, parts)) == [2, 3]                     # 5 videos over 2 ranks, contiguous shards
    merged = offline.merge_shards(parts)
    gold = np.load(os.path.join(ROOT, 'tests', 'golden', 'offline.npz'))['off2']
    assert list(merged) == G.NAMES
    assert np.array_equal(np.stack([merged[n] for n in G.NAMES]), gold)
