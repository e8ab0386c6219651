#!/usr/bin/env python
"""Roll out pre-extracted slots with a trained SlotFormer, with the reference's command line
(slotformer/video_prediction/rollout_clevrer_slots.py:101-128 / rollout_physion_slots.py):

    [torchrun --nproc-per-node G] python scripts/rollout_slots.py --params <slotformer_*_params.py> \
        --weight <ckpt.pth> --save_path <rollout_slots.pkl> [--obs-frames 128 --target-len 160]

``params.slots_root`` names the input ``{split: {video: float32[T, K, D]}}`` pickle (extract_slots.py output);
the result has the same structure with ``target_len`` frames per video.  One process per GPU: every rank rolls
out a contiguous shard of each split (no collective on the data path) and writes ``<save_path>.rank<r>``;
rank 0 merges the parts into ``save_path`` -- the reference pushes device_count() videos per step through
nn.DataParallel instead.
"""
import argparse
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'scripts'))


def main():
    from train import load_params
    from slotformer_b200 import offline
    from slotformer_b200.video_prediction import build_model
    ap = argparse.ArgumentParser(description='Rollout slots (sharded, one process per GPU)')
    ap.add_argument('--params', type=str, required=True)
    ap.add_argument('--weight', type=str, required=True, help='load weight')
    ap.add_argument('--save_path', type=str, required=True)
    ap.add_argument('--obs-frames', type=int, default=128)
    ap.add_argument('--target-len', type=int, default=160)
    ap.add_argument('--batch-videos', type=int, default=64)
    args = ap.parse_args()

    rank, world = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1))
    device = torch.device('cuda', int(os.environ.get('LOCAL_RANK', 0)))
    torch.cuda.set_device(device)
    params = load_params(args.params)
    model = build_model(params)
    model.load_state_dict(torch.load(args.weight, map_location='cpu')['state_dict'])
    model = model.to(device).eval()
    all_slots = offline.load_slots(params.slots_root)
    mine = {}
    for split, table in all_slots.items():
        names = offline.shard_names(list(table.keys()), rank, world)
        mine[split] = offline.rollout_video_slots(
            model.rollouter, {n: table[n] for n in names}, params.input_frames, params.frame_offset,
            obs_frames=args.obs_frames, target_len=args.target_len, batch_videos=args.batch_videos, device=device)
        print(f'[rank {rank}] {split}: {len(mine[split])}/{len(table)} videos', flush=True)
    part = f'{args.save_path}.rank{rank}'
    offline.dump_slots(mine, part)
    if rank == 0:
        parts = []
        for r in range(world):
            p = f'{args.save_path}.rank{r}'
            while not os.path.exists(p):          # files are the only rendezvous this job needs
                time.sleep(0.5)
            parts.append(offline.load_slots(p))
        merged = {s: offline.merge_shards([p.get(s, {}) for p in parts]) for s in all_slots}
        merged = {s: {n: merged[s][n] for n in all_slots[s]} for s in merged}      # original order
        offline.dump_slots(merged, args.save_path)
        for r in range(world):
            os.remove(f'{args.save_path}.rank{r}')
        print('wrote', args.save_path, {s: len(t) for s, t in merged.items()})


if __name__ == '__main__':
    main()
