#!/usr/bin/env python
"""Unified training entry point with the reference's command line (scripts/train.py:79-88):

    python scripts/train.py --task base_slots --params <..._params.py> [--weight ckpt] [--fp16] [--ddp]
                            [--cudnn] [--local_rank N] [--synthetic-steps 20]

Imports the task package (``slotformer_b200.<task>``), loads the params file exactly as the
reference does (``SlotFormerParams()`` from the file), builds the model with ``build_model`` and
runs a minimal trainer: Adam, linear-warm-up + cosine schedule (base_slots/method.py:71-87),
gradient clipping, ``<name>_loss_w`` loss weighting, ``{'state_dict': ...}`` checkpoints and -- with
--ddp -- ONE flattened NCCL all-reduce of the gradients per step.  Datasets are out of scope of this
repo (no data ships with it), so batches are synthetic tensors of the configured shapes; the loop is
the integration test of the drop-in model surface, not a data pipeline.
"""
import argparse
import importlib
import math
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def load_params(path):
    from slotformer_b200.compat import install_nerv_shim
    install_nerv_shim()
    sys.path.append(os.path.dirname(os.path.abspath(path)))
    name = os.path.basename(path)
    return importlib.import_module(name[:-3] if name.endswith('.py') else name).SlotFormerParams()


def synthetic_batch(task, params, model, device, gen):
    bs = max(2, int(params.get('train_batch_size', 4)) // 8)
    if task == 'base_slots':
        T = params.input_frames
        H, W = params.resolution
        return {'img': torch.rand((bs, T, 3, H, W), device=device, generator=gen) * 2 - 1}
    T = model.history_len + model.rollout_len
    return {'slots': torch.randn((bs, T, model.num_slots, model.slot_size), device=device, generator=gen)}


def lr_at(step, total, warmup, max_lr, min_lr=0.0):
    if step < warmup:
        return max_lr * (step + 1) / max(1, warmup)
    t = (step - warmup) / max(1, total - warmup)
    return min_lr + 0.5 * (max_lr - min_lr) * (1 + math.cos(math.pi * min(1.0, t)))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--task', type=str, required=True)
    ap.add_argument('--params', type=str, required=True)
    ap.add_argument('--weight', type=str, default='')
    ap.add_argument('--fp16', action='store_true')
    ap.add_argument('--ddp', action='store_true')
    ap.add_argument('--cudnn', action='store_true')
    ap.add_argument('--local_rank', '--local-rank', type=int, default=int(os.environ.get('LOCAL_RANK', 0)))
    ap.add_argument('--synthetic-steps', type=int, default=20)
    ap.add_argument('--out', type=str, default='')
    ap.add_argument('--synthetic-decoder', action='store_true',
                    help='video_prediction: mint the frozen SAVi decoder checkpoint from a fresh initialisation')
    ap.add_argument('--eval-mode-forward', action='store_true',
                    help='keep the model in eval() (no dropout): the rollout forward then runs on the kernel too')
    args = ap.parse_args()

    task = importlib.import_module(f'slotformer_b200.{args.task}')
    params = load_params(args.params)
    params.ddp = args.ddp
    if args.cudnn:
        torch.backends.cudnn.benchmark = True
    device = torch.device('cuda', args.local_rank)
    torch.cuda.set_device(device)
    if args.ddp:
        import torch.distributed as dist
        dist.init_process_group('nccl', device_id=device)
    if args.synthetic_decoder and not params.dec_dict.get('dec_ckp_path'):
        import tempfile
        from slotformer_b200.base_slots.models.savi import build_broadcast_decoder

        class _Dec(torch.nn.Module):
            pass
        torch.manual_seed(0)
        owner = _Dec()
        owner.dec_dict, owner.slot_size, owner.resolution = params.dec_dict, params.slot_dict['slot_size'], params.resolution
        build_broadcast_decoder(owner)
        path = os.path.join(tempfile.gettempdir(), f'sfb_synth_decoder_{os.getpid()}.pth')
        torch.save({'state_dict': owner.state_dict()}, path)
        params.dec_dict = dict(params.dec_dict, dec_ckp_path=path)
    torch.manual_seed(0)                                       # identical initial replicas on every rank
    model = task.build_model(params).to(device)
    if args.weight:
        model.load_state_dict(torch.load(args.weight, map_location='cpu')['state_dict'])
    model.train()
    if args.eval_mode_forward:
        model.eval()
    trainable = [p for p in model.parameters() if p.requires_grad]
    opt = torch.optim.Adam(trainable, lr=params.lr)
    scaler = torch.amp.GradScaler('cuda', enabled=args.fp16)
    total = args.synthetic_steps
    warmup = max(1, int(total * params.get('warmup_steps_pct', 0.025)))
    gen = torch.Generator(device=device).manual_seed(1 + args.local_rank)
    from slotformer_b200 import engine
    n_launch0 = engine.launch_count()
    ar_ms, step_ms, payload = [], [], 0
    for step in range(total):
        t_step = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t_step[0].record()
        for g in opt.param_groups:
            g['lr'] = lr_at(step, total, warmup, params.lr)
        batch = synthetic_batch(args.task, params, model, device, gen)
        with torch.autocast('cuda', dtype=torch.float16, enabled=args.fp16):
            out = model(batch)
            losses = model.calc_train_loss(batch, out)
            loss = sum(v * params.get(f'{k}_w', 1.0) for k, v in losses.items())
        opt.zero_grad(set_to_none=True)
        scaler.scale(loss).backward()
        scaler.unscale_(opt)
        if args.ddp:
            from slotformer_b200.parallel import allreduce_gradients
            t_ar = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t_ar[0].record()
            payload = allreduce_gradients(model) * 4               # ONE flattened NCCL all-reduce (fp32)
            t_ar[1].record()
        if params.get('clip_grad', -1) > 0:
            torch.nn.utils.clip_grad_norm_(trainable, params.clip_grad)
        scaler.step(opt)
        scaler.update()
        t_step[1].record()
        torch.cuda.synchronize(device)
        step_ms.append(t_step[0].elapsed_time(t_step[1]))
        if args.ddp:
            ar_ms.append(t_ar[0].elapsed_time(t_ar[1]))
        if args.local_rank == 0 and (step % 5 == 0 or step == total - 1):
            print(f'step {step:4d}  loss {loss.item():.5f}  ' +
                  '  '.join(f'{k} {v.item():.5f}' for k, v in losses.items()), flush=True)
    if args.local_rank == 0:
        import json
        import statistics
        summary = {'task': args.task, 'params': os.path.basename(args.params), 'steps': total,
                   'world_size': int(os.environ.get('WORLD_SIZE', 1)),
                   'ms_per_step_median': statistics.median(step_ms[2:] or step_ms),
                   'kernel_launches_per_step': (engine.launch_count() - n_launch0) / total}
        if ar_ms:
            summary.update(allreduce_payload_bytes=payload, allreduce_ms_median=statistics.median(ar_ms[2:] or ar_ms),
                           allreduce_ms_min=min(ar_ms[2:] or ar_ms), allreduce_backend='nccl')
        print('TRAIN_SUMMARY ' + json.dumps(summary), flush=True)
    if args.out and args.local_rank == 0:
        torch.save({'state_dict': model.state_dict(), 'it': total}, args.out)
    if args.ddp:
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
