#!/bin/bash
mkdir -p gpurun_out
# launch list of the bench command (per-launch times are cold-cache and serialised: shares only)
timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 60 -c 16 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu-baseline > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/launches_bench.csv')) if len(r)>5 and r[0].isdigit()]
for r in rows: print(r[0], r[4][:60], r[-3], r[-2], r[-1])
PY
# train.py integration smoke (synthetic batches) if a reference-style params file is available
cat > /tmp/savi_small_params.py <<'PY2'
from nerv.training import BaseParams
class SlotFormerParams(BaseParams):
    project = 'x'
    lr = 1e-4
    warmup_steps_pct = 0.05
    clip_grad = 0.05
    train_batch_size = 16
    model = 'StoSAVi'
    resolution = (64, 64)
    input_frames = 3
    slot_dict = dict(num_slots=5, slot_size=128, slot_mlp_size=256, num_iterations=2, kernel_mlp=True)
    enc_dict = dict(enc_channels=(3, 64, 64, 64, 64), enc_ks=5, enc_out_channels=128, enc_norm='')
    dec_dict = dict(dec_channels=(128, 64, 64, 64, 64), dec_resolution=(8, 8), dec_ks=5, dec_norm='')
    pred_dict = dict(pred_type='transformer', pred_rnn=True, pred_norm_first=True, pred_num_layers=2, pred_num_heads=4, pred_ffn_dim=512, pred_sg_every=None)
    loss_dict = dict(use_post_recon_loss=True, kld_method='none')
    post_recon_loss_w = 1.
    kld_loss_w = 1e-4
PY2
timeout 200 python scripts/train.py --task base_slots --params /tmp/savi_small_params.py --synthetic-steps 11 2>&1 | tail -5
