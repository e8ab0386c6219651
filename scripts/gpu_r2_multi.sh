#!/bin/bash
# 2-GPU job: NCCL sharding bit-identity test, DDP training runs (the one collective), sharded offline job, bench N=2
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_multirank.py tests/test_offline.py -q -m gpu --tb=short > gpurun_out/t_multi.log 2>&1; echo "multirank tests rc=$?"; tail -3 gpurun_out/t_multi.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533"
timeout 300 $TR scripts/train.py --task video_prediction --params scripts/configs/slotformer_synth_params.py --ddp --synthetic-decoder --eval-mode-forward --synthetic-steps 20 > gpurun_out/train_ddp2_slotformer.log 2>&1; echo "train slotformer rc=$?"; grep TRAIN_SUMMARY gpurun_out/train_ddp2_slotformer.log
timeout 300 $TR scripts/train.py --task base_slots --params scripts/configs/savi_synth_params.py --ddp --synthetic-steps 20 > gpurun_out/train_ddp2_savi.log 2>&1; echo "train savi rc=$?"; grep TRAIN_SUMMARY gpurun_out/train_ddp2_savi.log
timeout 400 $TR bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_r2_n2.json 2> gpurun_out/bench_r2_n2.err; echo "bench n2 rc=$?"; cat gpurun_out/bench_r2_n2.json | cut -c1-400
