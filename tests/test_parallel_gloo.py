"""world_size-2 gloo test (CPU) of the multi-GPU host logic: clip sharding with no data-path
collective, result gathering, and the single flattened gradient all-reduce."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import cases
from oracle import slot_oracle as O


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    from slotformer_b200 import parallel
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        # --- inference: every rank runs the path on its own clips, nothing is exchanged ---
        c, w, feats, slots = cases.sa_case('sa_tiny')
        B = 5
        rs = np.random.RandomState(3)
        feats = rs.standard_normal((B, 64, c['C'])).astype(np.float32)
        slots = rs.standard_normal((B, c['K'], c['D'])).astype(np.float32)
        f_loc = parallel.shard_clips(torch.from_numpy(feats))
        s_loc = parallel.shard_clips(torch.from_numpy(slots))
        out_loc = torch.from_numpy(O.slot_attention(f_loc.numpy(), s_loc.numpy(), w, 1, dtype=np.float32))
        full = parallel.gather_clips(out_loc, B)
        ref = O.slot_attention(feats, slots, w, 1, dtype=np.float32)
        ok_gather = bool(np.allclose(full.numpy(), ref, rtol=0, atol=1e-6))
        # --- training: one flattened all-reduce == single-process gradient of the mean loss ---
        torch.manual_seed(0)
        net = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.ReLU(), torch.nn.Linear(16, 4))
        x = torch.from_numpy(rs.standard_normal((6, 8)).astype(np.float32))
        y = torch.from_numpy(rs.standard_normal((6, 4)).astype(np.float32))
        xs, ys = parallel.shard_clips(x), parallel.shard_clips(y)
        torch.nn.functional.mse_loss(net(xs), ys).backward()
        n = parallel.allreduce_gradients(net)
        got = torch.cat([p.grad.reshape(-1) for p in net.parameters()])
        net.zero_grad()
        torch.nn.functional.mse_loss(net(x), y).backward()
        want = torch.cat([p.grad.reshape(-1) for p in net.parameters()])
        ok_grad = bool(torch.allclose(got, want, atol=1e-6)) and n == got.numel()
        q.put((rank, ok_gather, ok_grad, tuple(f_loc.shape)))
    finally:
        dist.destroy_process_group()


def test_shard_bounds_cover_everything():
    from slotformer_b200.parallel import shard_bounds
    for n in (0, 1, 7, 64, 65):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_bounds(4, 2, 2)


def test_two_rank_sharding_and_gradient_allreduce():
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [r[0] for r in results] == [0, 1]
    assert all(r[1] and r[2] for r in results), results
    assert results[0][3][0] == 3 and results[1][3][0] == 2      # 5 clips -> 3 + 2
