"""Multi-GPU plumbing: one process per GPU, clips sharded across ranks.

Both hot paths are independent per clip, so inference needs NO data-path
collective: every rank runs the kernels on its own contiguous shard; ``gather_clips`` is only for
collecting results on every rank (e.g. rank-0 file writing).  Training adds exactly one gradient
all-reduce per step over a single flattened buffer (``allreduce_gradients``), replacing the DDP
wrap the reference gets from nerv (scripts/train.py:65-76, sbatch_run.sh:36-42)."""
import os

import torch
import torch.distributed as dist


def bind_to_gpu_numa(device_index):
    """Pin this process (and the threads it starts later) to the CPUs NVML reports as local to the GPU, so that pinned
    host buffers allocated afterwards are first-touched on the GPU's own NUMA node and H2D / D2H copies do not cross
    the socket interconnect (8 ranks feeding 8 GPUs from one node is host-bandwidth bound).
    Returns the CPU set, or None when NVML or the affinity call is unavailable (nothing is changed then)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        visible = os.environ.get('CUDA_VISIBLE_DEVICES')
        idx = int(device_index)
        if visible:
            ids = [v.strip() for v in visible.split(',') if v.strip()]
            if idx < len(ids) and ids[idx].isdigit():
                idx = int(ids[idx])
        handle = pynvml.nvmlDeviceGetHandleByIndex(idx)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(handle, (ncpu + 63) // 64)
        cpus = {64 * i + b for i, w in enumerate(words) for b in range(64) if (int(w) >> b) & 1}
        cpus &= set(os.sched_getaffinity(0))
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return sorted(cpus)
    except Exception:
        return None


def shard_bounds(n_items, rank, world_size):
    """Contiguous [start, stop) of rank's shard; sizes differ by at most one (earlier ranks larger)."""
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError(f'bad rank {rank} / world size {world_size}')
    base, extra = divmod(n_items, world_size)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def shard_clips(tensor, rank=None, world_size=None):
    """This rank's contiguous slice of a [B, ...] batch of clips."""
    rank = dist.get_rank() if rank is None else rank
    world_size = dist.get_world_size() if world_size is None else world_size
    lo, hi = shard_bounds(tensor.shape[0], rank, world_size)
    return tensor[lo:hi]


def gather_clips(local, n_items, group=None):
    """Inverse of shard_clips: every rank receives the full [n_items, ...] tensor."""
    world = dist.get_world_size(group)
    sizes = [shard_bounds(n_items, r, world) for r in range(world)]
    longest = max(hi - lo for lo, hi in sizes)
    pad = local.new_zeros((longest,) + tuple(local.shape[1:]))
    pad[:local.shape[0]] = local
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad, group=group)
    return torch.cat([p[:hi - lo] for p, (lo, hi) in zip(parts, sizes)], dim=0)


def allreduce_gradients(module, group=None, average=True):
    """ONE all-reduce over all gradients of ``module`` (flattened), in place.  Payloads here are
    3-26 MB: latency bound, so a single call beats per-tensor buckets."""
    grads = [p.grad for p in module.parameters() if p.grad is not None]
    if not grads:
        return 0
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    if average:
        flat /= dist.get_world_size(group)
    off = 0
    for g in grads:
        n = g.numel()
        g.copy_(flat[off:off + n].view_as(g))
        off += n
    return flat.numel()
