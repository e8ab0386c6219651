#!/usr/bin/env python
"""Extract slots from videos with a trained SAVi / STEVE model, with the reference's command lines
(slotformer/base_slots/extract_slots.py:135-158, extract_phyre_slots.py:79-113):

    [torchrun --nproc-per-node G] python scripts/extract_slots.py --params <savi_*_params.py> --weight <ckpt.pth> \
        --videos <videos.pkl | dir of <name>.npy> --save_path <slots.pkl>
    python scripts/extract_slots.py --phyre --params ... --weight ... --videos <dir of {idx:06d}.npy> \
        --save_path <root> --split 3 --total_split 8 [--bs 8]

The reference reads its datasets (OBJ3D / CLEVRER / Physion / PHYRE simulators), which are outside the hot-path
scope and absent here; the video source is therefore explicit: a pickle ``{split: {name: float[T,3,H,W] in [-1,1]}}``
or a directory of ``.npy`` files of that shape.  Output formats are the reference's: ``{split: {name: float32[T,K,D]}}``
(.pkl) or, with --phyre, one ``{idx:06d}.npy`` per sample cut to ``vid_len`` (a sibling ``{idx:06d}.len`` text file,
default = all frames).  The model runs with ``testing = True`` (slots only).  Video mode: one process per GPU, each rank
extracts a contiguous shard and rank 0 merges (the reference pushes one video per GPU through nn.DataParallel).
PHYRE mode: ``--split i --total_split n`` is one independent job of n, resumable, exactly as
scripts/parallel_phyre.sh launches them.
"""
import argparse
import os
import pickle
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'scripts'))


def load_videos(path):
    """-> {split: {name: loader()}}"""
    if os.path.isdir(path):
        names = sorted(f[:-4] for f in os.listdir(path) if f.endswith('.npy'))
        return {'val': {n: (lambda n=n: np.load(os.path.join(path, n + '.npy'))) for n in names}}
    with open(path, 'rb') as f:
        table = pickle.load(f)
    return {s: {n: (lambda v=v: v) for n, v in t.items()} for s, t in table.items()}


def main():
    from train import load_params
    from slotformer_b200 import offline
    from slotformer_b200.base_slots import build_model
    ap = argparse.ArgumentParser(description='Extract slots from videos (sharded, one process per GPU)')
    ap.add_argument('--params', type=str, required=True)
    ap.add_argument('--weight', type=str, required=True, help='pretrained model weight')
    ap.add_argument('--videos', type=str, required=True)
    ap.add_argument('--save_path', type=str, required=True)
    ap.add_argument('--phyre', action='store_true')
    ap.add_argument('--split', type=int, default=-1)
    ap.add_argument('--total_split', type=int, default=10)
    ap.add_argument('--bs', type=int, default=8)
    args = ap.parse_args()

    rank, world = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1))
    device = torch.device('cuda', int(os.environ.get('LOCAL_RANK', 0)))
    torch.cuda.set_device(device)
    params = load_params(args.params)
    model = build_model(params)
    model.load_state_dict(torch.load(args.weight, map_location='cpu')['state_dict'])
    model.testing = True
    model = model.to(device).eval()
    slot_key = 'post_slots' if params.model == 'StoSAVi' else 'slots'

    if args.phyre:
        vids = sorted(f for f in os.listdir(args.videos) if f.endswith('.npy'))

        def get_sample(i):
            v = np.load(os.path.join(args.videos, f'{i:06d}.npy'))
            lp = os.path.join(args.videos, f'{i:06d}.len')
            return v, (int(open(lp).read()) if os.path.exists(lp) else v.shape[0])

        root = os.path.join(args.save_path, 'slots', os.path.basename(args.params))
        done = offline.extract_phyre_job(model, get_sample, len(vids), root, args.split, args.total_split, args.bs,
                                         device=device, slot_key=slot_key)
        print(f'split {args.split}/{args.total_split}: wrote {len(done)} samples under {root}')
        return

    videos = load_videos(args.videos)
    mine = {}
    for split, table in videos.items():
        names = offline.shard_names(list(table.keys()), rank, world)
        mine[split] = offline.extract_video_slots(model, lambda n: torch.as_tensor(table[n]()), names,
                                                  batch_videos=args.bs, device=device, slot_key=slot_key)
        print(f'[rank {rank}] {split}: {len(mine[split])}/{len(table)} videos', flush=True)
    offline.dump_slots(mine, f'{args.save_path}.rank{rank}')
    if rank == 0:
        parts = []
        for r in range(world):
            p = f'{args.save_path}.rank{r}'
            while not os.path.exists(p):          # files are the only rendezvous this job needs
                time.sleep(0.5)
            parts.append(offline.load_slots(p))
        merged = {s: offline.merge_shards([p.get(s, {}) for p in parts]) for s in videos}
        merged = {s: {n: merged[s][n] for n in videos[s]} for s in merged}
        offline.dump_slots(merged, args.save_path)
        for r in range(world):
            os.remove(f'{args.save_path}.rank{r}')
        print('wrote', args.save_path, {s: len(t) for s, t in merged.items()})


if __name__ == '__main__':
    main()
