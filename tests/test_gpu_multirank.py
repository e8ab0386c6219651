"""Two ranks over NCCL on two GPUs (skipped on a one-GPU box): the clip-sharded kernels give bit-identical slots and
predictions to a single rank processing the whole batch -- clips are independent, so sharding may not change a bit."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import cases
from helpers import golden

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    from helpers import ro_module, sa_module
    from slotformer_b200 import parallel
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device('cuda', rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=dev)
    try:
        c, w, _, _ = cases.sa_case('sa_cfg2')
        sa = sa_module(c, w, dev, mask=False)
        rc, rw, _ = cases.ro_case('ro_cfg2')
        ro = ro_module(rc, rw, dev, enc_t_pe=golden('ro_cfg2')['enc_t_pe'])
        B, T, K, D = 10, 6, 6, 128                     # 10 clips -> 5 + 5; frames = clips * T
        gen = torch.Generator(device=dev).manual_seed(17)              # same stream on every rank
        feats = torch.randn((B * T, 4096, 128), device=dev, generator=gen)
        init = torch.randn((B * T, K, D), device=dev, generator=gen)
        with torch.no_grad():
            lo, hi = parallel.shard_bounds(B, rank, world)
            s_loc = sa(feats[lo * T:hi * T], init[lo * T:hi * T])
            p_loc = ro(s_loc.view(hi - lo, T, K, D), 10)
            slots = parallel.gather_clips(s_loc.view(hi - lo, T, K, D), B)          # NCCL all_gather
            pred = parallel.gather_clips(p_loc, B)
            s_all = sa(feats, init)
            p_all = ro(s_all.view(B, T, K, D), 10)
        ok = bool(torch.equal(slots, s_all.view(B, T, K, D))) and bool(torch.equal(pred, p_all))
        q.put((rank, ok, dist.get_backend()))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs two GPUs')
def test_two_rank_nccl_sharding_is_bit_identical_to_single_rank():
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = sorted(q.get(timeout=300) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [r[0] for r in results] == [0, 1]
    assert all(r[1] and r[2] == 'nccl' for r in results), results
