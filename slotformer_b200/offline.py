"""Offline drivers either side of the two hot paths.

The reference extracts slots and rolls them out with single-process scripts that push
``torch.cuda.device_count()`` videos per step through ``nn.DataParallel``
(slotformer/base_slots/extract_slots.py:18-38, extract_phyre_slots.py:64-76,
slotformer/video_prediction/rollout_clevrer_slots.py:19-66, rollout_physion_slots.py:22-62).
Here the same jobs are sharded one-process-per-GPU (clips are independent: no data-path collective), each rank
feeds whole batches of videos to the persistent kernels, and the on-disk formats are kept byte-compatible:

  * ``slots.pkl`` / ``rollout_slots.pkl``: pickle of ``{split: {video_name: float32[T, K, D]}}``
    (extract_slots.py:58-76, rollout_clevrer_slots.py:69-98);
  * PHYRE: one ``{index:06d}.npy`` per sample, cut to the real video length (extract_phyre_slots.py:70-76).

Everything numeric is done by the model passed in (our modules -> the sm_100a kernels); this file is index
arithmetic, batching, sharding and file formats only.
"""
import os
import pickle

import numpy as np
import torch

from .parallel import shard_bounds


# ----------------------------------------------------------------------------------------------
# file formats
# ----------------------------------------------------------------------------------------------
def dump_slots(slots, path):
    """``{split: {name: float32[T, K, D]}}`` -> pickle file (what nerv.utils.dump_obj writes for a .pkl path)."""
    for split, table in slots.items():
        for name, arr in table.items():
            if not (isinstance(arr, np.ndarray) and arr.dtype == np.float32 and arr.ndim == 3):
                raise ValueError(f'{split}/{name}: expected float32 [T, K, D], got {type(arr)} '
                                 f'{getattr(arr, "dtype", None)} {getattr(arr, "shape", None)}')
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    tmp = path + '.tmp'
    with open(tmp, 'wb') as f:
        pickle.dump(slots, f, protocol=pickle.HIGHEST_PROTOCOL)
    os.replace(tmp, path)


def load_slots(path):
    with open(path, 'rb') as f:
        return pickle.load(f)


def save_phyre_sample(save_root, index, slots, vid_len):
    """extract_phyre_slots.py:70-76: ``{index:06d}.npy`` holding slots[:vid_len]."""
    os.makedirs(save_root, exist_ok=True)
    np.save(os.path.join(save_root, f'{int(index):06d}.npy'), np.ascontiguousarray(slots[:int(vid_len)], dtype=np.float32))


# ----------------------------------------------------------------------------------------------
# sharding helpers (one process per GPU; no collective on the data path)
# ----------------------------------------------------------------------------------------------
def shard_names(names, rank, world_size):
    """Contiguous shard of a list of video names (same rule as parallel.shard_clips)."""
    lo, hi = shard_bounds(len(names), rank, world_size)
    return list(names[lo:hi])


def merge_shards(parts):
    """Per-rank ``{name: array}`` tables -> one table; a name may come from one rank only."""
    out = {}
    for part in parts:
        dup = set(out) & set(part)
        if dup:
            raise ValueError(f'videos present in more than one shard: {sorted(dup)[:3]}')
        out.update(part)
    return out


# ----------------------------------------------------------------------------------------------
# frame-offset rollout (index arithmetic of rollout_clevrer_slots.py:44-58 / rollout_physion_slots.py:40-54)
# ----------------------------------------------------------------------------------------------
def offset_starts(obs_frames, history_len, frame_offset):
    """First observed frame fed to the model for every offset (rollout_clevrer_slots.py:48)."""
    starts = [obs_frames - history_len * frame_offset + off for off in range(frame_offset)]
    if starts[0] < 0:
        raise ValueError(f'{obs_frames} observed frames cannot hold {history_len} history frames at offset {frame_offset}')
    return starts


def interleave_predictions(per_offset, n_future):
    """``per_offset[o]``: [B, steps_o, K, D] predictions of the sub-sequence o, o+f, o+2f, ...  -> [B, n_future, K, D]
    with future frame i taken from offset i % f, step i // f (rollout_clevrer_slots.py:55-58)."""
    f = len(per_offset)
    return torch.stack([per_offset[i % f][:, i // f] for i in range(n_future)], dim=1)


@torch.no_grad()
def rollout_video_slots(rollouter, pre_slots, history_len, frame_offset, obs_frames, target_len, batch_videos=64,
                        device='cuda:0'):
    """Observed slots of every video -> observed + rolled-out slots.

    ``rollouter(x [B, history_len, K, D], pred_len) -> [B, pred_len, K, D]`` (SlotRollouter.forward; for a whole
    SlotFormer model pass ``model.rollouter``).  ``pre_slots``: ``{name: float32[T >= obs_frames, K, D]}``.
    A model trained with frame offset f sees every f-th frame, so the future is rolled out as f interleaved
    sub-sequences, exactly as the reference does; all f sub-sequences of a batch go through ONE kernel launch
    (they are independent clips) when they need the same number of steps.
    Returns ``{name: float32[target_len, K, D]}``.
    """
    names = list(pre_slots.keys())
    n_future = target_len - obs_frames
    starts = offset_starts(obs_frames, history_len, frame_offset)
    # steps needed by offset o: future frames i with i % f == o  ->  indices o, o+f, ...
    steps = [len(range(o, n_future, frame_offset)) for o in range(frame_offset)]
    out = {}
    for b0 in range(0, len(names), batch_videos):
        chunk = names[b0:b0 + batch_videos]
        obs = torch.from_numpy(np.stack([np.asarray(pre_slots[n][:obs_frames], dtype=np.float32) for n in chunk]))
        obs = obs.to(device)
        hist = [obs[:, s::frame_offset][:, :history_len] for s in starts]          # f x [B, history_len, K, D]
        if len(set(steps)) == 1:                                                    # one launch for all offsets
            pred = rollouter(torch.cat(hist, dim=0).contiguous(), steps[0])
            per_offset = list(pred.split(len(chunk), dim=0))
        else:
            per_offset = [rollouter(h.contiguous(), max(st, 1))[:, :st] for h, st in zip(hist, steps)]
        fut = interleave_predictions(per_offset, n_future)
        full = torch.cat([obs, fut], dim=1).cpu().numpy()
        assert full.shape[1] == target_len
        for i, n in enumerate(chunk):
            out[n] = full[i]
    return out


@torch.no_grad()
def extract_video_slots(model, get_video, names, batch_videos=8, device='cuda:0', slot_key='post_slots'):
    """``get_video(name) -> float tensor [T, 3, H, W]``; runs ``model({'img': [B, T, 3, H, W]})[slot_key]`` over
    batches of equally long videos (extract_slots.py:18-38) and returns ``{name: float32[T, K, D]}``."""
    model.eval()
    out = {}
    pending = []

    def flush():
        if not pending:
            return
        vids = torch.stack([v for _, v in pending]).float().to(device)
        slots = model({'img': vids})[slot_key].detach().float().cpu().numpy()
        for (n, _), s in zip(pending, slots):
            out[n] = np.ascontiguousarray(s, dtype=np.float32)
        pending.clear()

    for n in names:
        v = get_video(n)
        if pending and (v.shape != pending[0][1].shape or len(pending) == batch_videos):
            flush()
        pending.append((n, v))
    flush()
    return out


# ----------------------------------------------------------------------------------------------
# PHYRE extraction: job-level sharding + resume (extract_phyre_slots.py:41-76, scripts/parallel_phyre.sh:22-28)
# ----------------------------------------------------------------------------------------------
def phyre_split_range(total_num, split, total_split):
    """Sample range [start, end) of job ``split`` of ``total_split`` (extract_phyre_slots.py:41-44): equal shares of
    ``total_num // total_split``, the last job takes the remainder; ``split == -1`` is the whole set."""
    if split == -1:
        return 0, total_num
    if not 0 <= split < total_split:
        raise ValueError(f'split {split} outside [0, {total_split})')
    share = total_num // total_split
    return share * split, (share * (split + 1) if split < total_split - 1 else total_num)


def phyre_resume_start(save_root, start_idx, end_idx):
    """First sample this job still has to do (extract_phyre_slots.py:45-53): scan for the first missing
    ``{idx:06d}.npy`` and go back by one, "in case the last file is corrupted".  The reference clamps at 0
    (`max(idx - 1, 0)`), so a fresh job re-writes the last file of its neighbour; here the clamp is the job's own
    start, which keeps concurrently running jobs from touching each other's files."""
    idx = start_idx
    for idx in range(start_idx, end_idx):
        if not os.path.exists(os.path.join(save_root, f'{idx:06d}.npy')):
            break
    return max(idx - 1, start_idx)


@torch.no_grad()
def extract_phyre_job(model, get_sample, total_num, save_root, split=-1, total_split=1, batch_size=8,
                      device='cuda:0', slot_key='post_slots', resume=True):
    """One shard of the PHYRE extraction job.  ``get_sample(idx) -> (video float [T, 3, H, W], vid_len)``; every
    sample's slots are saved to ``save_root/{idx:06d}.npy`` cut to its real length.  Jobs are independent (one per GPU
    / Slurm job, no collective); a restarted job skips what it already wrote.  Returns the indices written."""
    model.eval()
    os.makedirs(save_root, exist_ok=True)
    start, end = phyre_split_range(total_num, split, total_split)
    if split != -1 and resume:
        start = phyre_resume_start(save_root, start, end)
    written = []
    for b0 in range(start, end, batch_size):
        idxs = list(range(b0, min(b0 + batch_size, end)))
        samples = [get_sample(i) for i in idxs]
        vids = torch.stack([torch.as_tensor(v) for v, _ in samples]).float().to(device)
        slots = model({'img': vids})[slot_key].detach().float().cpu().numpy()
        for i, (_, vid_len), s in zip(idxs, samples, slots):
            save_phyre_sample(save_root, i, s, vid_len)
            written.append(i)
    return written
