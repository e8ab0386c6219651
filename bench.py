#!/usr/bin/env python
"""Headline benchmark: video-clip frames/sec (slot extraction fwd + rollout) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload = BASELINE.json configs[1] ("OBJ3D SlotFormer rollout, B=64, 128x128, K=6, T 6->10"):
one step = Slot Attention over the B*T_in = 384 burn-in frames' CNN feature grids
([384, 4096, 128] fp32, 2 iterations, K=6 slots, D=128) followed by the 10-step
autoregressive slot-Transformer rollout (d=128, 4 layers, 8 heads, F=512) of the B=64 clips.
frames/s = B * (T_in + T_out) / time.  Synthetic inputs, seeded weights (tests/golden/cases.py).

  value   : inputs already resident in HBM, CUDA-event timed, max over ranks (weak scaling:
            every rank runs its own B=64 clips; the path has no data-path collective).
  e2e     : the same clips end to end through the model API a user calls, HOST (pinned) buffers, every copy
            inside the timed region: images [64, 6, 3, 128, 128] fp32 (75 MB) -> StoSAVi.encode (cuDNN CNN, the fused
            encoder tail sfb_enc_tail_forward, per frame sfb_transition_forward + sfb_sa_forward, replayed as one
            CUDA graph) -> SlotRollouter (sfb_rollout_forward) -> slots + predictions back on the host.
            e2e.cnn_output_route enters one step later (host input = the CNN encoder's output, the tensor
            _get_encoder_out consumes at savi.py:367, 404 MB per step) and e2e.feature_grid_route at the operator
            boundary itself (the fp32 feature grid, 806 MB per step): both are PCIe-bound on one GPU and bound by the
            host's aggregate H2D bandwidth when 8 ranks share it.
  variants: SURVEY 8(d) (E) / (E+D): the model-level step with device-resident images, without / with the decoder.
  roofline: the Slot Attention kernel against the measured HBM peak (algorithmic bytes =
            N*C*4 + 2*K*D*4 per frame).
  cpu_baseline / --impl reference: the UNMODIFIED reference modules (oracle/_ref, collected by
            oracle/build_ref.py) on the host cores, full 64-clip steps (kind "reference").
  gpu_eager_baseline: the same unmodified modules in PyTorch eager on the same GPU (fp32 and
            bf16 autocast) -- the ">= 10x reference single-GPU forward" denominator.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, 'tests', 'golden')):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import numpy as np  # noqa: E402

# ---- workload (BASELINE.json configs[1]) ------------------------------------------------
WL = dict(B=64, T_in=6, T_out=10, K=6, N=4096, C=128, D=128, Dm=256, iters=2,
          d=128, layers=4, heads=8, F=512)
METRIC = 'video_clip_frames_per_sec'
UNIT = 'frames/s'
WORKLOAD = ('OBJ3D SlotFormer rollout, B=64, 128x128 (64x64 feature grid), K=6, T 6->10: '
            'SlotAttention(384 frames x 4096 x 128, 2 it) + SlotRollouter(d=128, 4L, 8H, F=512, 10 steps)')


# dram__bytes_read.sum + dram__bytes_write.sum over the kernels of ONE sfb_sa_forward call on this
# workload, from the ncu --set full capture summarised in profiles/ (None until captured)
SA_TRAFFIC_BYTES = 1558e6
SA_TRAFFIC_SOURCE = 'profiles/r2b_sa_per_kernel_tc_vs_mma.txt (ncu dram__bytes_read+write over the 5 SA launches of one step)'


def sa_bytes_per_frame():
    return WL['N'] * WL['C'] * 4 + 2 * WL['K'] * WL['D'] * 4


def ro_flops_total():
    L = WL['T_in'] * WL['K']
    d, F, Ds, K = WL['d'], WL['F'], WL['D'], WL['K']
    per_step = 2 * L * Ds * d + WL['layers'] * (8 * L * d * d + 4 * L * L * d + 4 * L * d * F) + 2 * K * d * Ds
    return per_step * WL['T_out'] * WL['B']


def frames_per_step():
    return WL['B'] * (WL['T_in'] + WL['T_out'])


def peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return dict(hbm=float(p['hbm_gbs']), tf=float(p.get('bf16_tflops_sustained', p['bf16_tflops'])),
                    src='measured (MEASURED_PEAKS.json)')
    return dict(hbm=6650.0, tf=1400.0, src='fallback (B200_PROFILING.md)')


def make_weights():
    import cases
    sa_w = cases.make_sa_weights(WL['C'], WL['D'], WL['Dm'], seed=13)
    ro_w = cases.make_ro_weights(WL['D'], WL['d'], WL['F'], WL['layers'], seed=22)
    return sa_w, ro_w


def make_tail_weights(seed=31):
    """Encoder tail of StoSAVi (savi.py:126-133, utils.py:52-63): SoftPositionEmbed dense(4 -> 64), LayerNorm(64),
    Linear(64 -> C), ReLU, Linear(C -> C); seeded like the other synthetic weights."""
    rs = np.random.RandomState(seed)
    C = WL['C']

    def lin(o, i):
        a = np.sqrt(3.0 / i)
        return rs.uniform(-a, a, size=(o, i)).astype(np.float32)

    return {'encoder_pos_embedding.dense.weight': lin(64, 4),
            'encoder_pos_embedding.dense.bias': (0.1 * rs.standard_normal(64)).astype(np.float32),
            'encoder_out_layer.0.weight': (1.0 + 0.1 * rs.standard_normal(64)).astype(np.float32),
            'encoder_out_layer.0.bias': (0.1 * rs.standard_normal(64)).astype(np.float32),
            'encoder_out_layer.1.weight': lin(C, 64),
            'encoder_out_layer.1.bias': (0.1 * rs.standard_normal(C)).astype(np.float32),
            'encoder_out_layer.3.weight': lin(C, C),
            'encoder_out_layer.3.bias': (0.1 * rs.standard_normal(C)).astype(np.float32)}


# ---- clocks sampler (NVML) --------------------------------------------------------------
class ClockSampler(threading.Thread):
    REASONS = {0x1: 'gpu_idle', 0x2: 'applications_clocks_setting', 0x4: 'sw_power_cap',
               0x8: 'hw_slowdown', 0x10: 'sync_boost', 0x20: 'sw_thermal_slowdown',
               0x40: 'hw_thermal_slowdown', 0x80: 'hw_power_brake_slowdown',
               0x100: 'display_clock_setting'}

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.stop_flag = threading.Event()
        self.sm = []
        self.mask = 0
        self.max_mhz = None
        self.err = None

    def run(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            while not self.stop_flag.is_set():
                self.sm.append(int(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)))
                try:
                    self.mask |= int(pynvml.nvmlDeviceGetCurrentClocksEventReasons(h))
                except Exception:
                    self.mask |= int(pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h))
                time.sleep(0.002)
        except Exception as e:  # noqa: BLE001
            self.err = repr(e)

    def result(self):
        self.stop_flag.set()
        self.join(timeout=2)
        reasons = [n for b, n in self.REASONS.items() if (self.mask & b) and n != 'gpu_idle']
        out = {'sm_mhz': int(np.median(self.sm)) if self.sm else None, 'sm_max_mhz': self.max_mhz,
               'reasons': reasons, 'samples': len(self.sm)}
        if self.err:
            out['error'] = self.err
        return out


# ---- reference arms: the UNMODIFIED reference modules (oracle/_ref, see oracle/build_ref.py) ------------
def cpu_threads():
    return os.cpu_count() or 1


def ref_modules(device, dtype=None):
    """Reference SlotAttention (savi.py:16-110) and SlotRollouter (slotformer.py:48-134), unmodified, with the
    bench's seeded weights."""
    import torch
    from oracle.build_ref import import_ref
    R = import_ref()
    sa_w, ro_w = make_weights()
    sa = R['SlotAttention'](in_features=WL['C'], num_iterations=WL['iters'], num_slots=WL['K'], slot_size=WL['D'],
                            mlp_hidden_size=WL['Dm'])
    sa.load_state_dict({k: torch.from_numpy(v) for k, v in sa_w.items()}, strict=True)
    ro = R['SlotRollouter'](num_slots=WL['K'], slot_size=WL['D'], history_len=WL['T_in'], d_model=WL['d'],
                            num_layers=WL['layers'], num_heads=WL['heads'], ffn_dim=WL['F'])
    missing, unexpected = ro.load_state_dict({k: torch.from_numpy(v) for k, v in ro_w.items()}, strict=False)
    assert not unexpected and set(missing) <= {'enc_t_pe'}
    return sa.to(device).eval(), ro.to(device).eval()


def ref_step(sa, ro, feats, init):
    """One step of the workload through the reference's own code path: Slot Attention frame by frame as
    StoSAVi.encode calls it (savi.py:393-410: [B, N, C] per time step), then SlotRollouter.forward.
    feats [B, T_in, N, C], init [T_in, B, K, D] (the per-frame initial slots the predictor would hand over, each
    contiguous like kernel_dist_layer's output) -> (slots [B, T_in, K, D], pred [B, T_out, K, D])."""
    import torch
    slots = torch.stack([sa(feats[:, t], init[t]) for t in range(feats.shape[1])], dim=1)
    return slots, ro(slots, WL['T_out'])


def ref_inputs(device, seed=1):
    import torch
    gen = torch.Generator().manual_seed(seed)
    B, T = WL['B'], WL['T_in']
    feats = torch.randn((B, T, WL['N'], WL['C']), generator=gen)
    feats.mul_(0.5 + 1.5 * torch.rand((B, T, WL['N'], 1), generator=gen))
    init = torch.randn((T, B, WL['K'], WL['D']), generator=gen)
    return feats.to(device), init.to(device)


def cpu_reference_times(steps, warmup):
    """Full 64-clip steps of the unmodified reference on the host cores (fp32, eval, no_grad)."""
    import torch
    torch.set_num_threads(cpu_threads())          # torchrun exports OMP_NUM_THREADS=1; the CPU arm may use every core
    sa, ro = ref_modules('cpu')
    feats, init = ref_inputs('cpu')
    ts = []
    with torch.no_grad():
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            ref_step(sa, ro, feats, init)
            if i >= warmup:
                ts.append(time.perf_counter() - t0)
    return ts


def cpu_baseline(repeats=2):
    ts = cpu_reference_times(repeats, 1)
    t = float(np.median(ts))
    return {'value': frames_per_step() / t, 'unit': UNIT, 'cores': cpu_threads(), 'kind': 'reference',
            'sample': f'the full step ({WL["B"]} clips: {WL["B"] * WL["T_in"]} frames of Slot Attention + {WL["B"]}-clip rollout) '
                      f'through the unmodified reference modules (oracle/_ref), torch {_torch_version()} CPU fp32, '
                      f'median of {repeats} after 1 warm-up, {t:.2f} s per step'}


def _torch_version():
    import torch
    return torch.__version__


def gpu_eager_baseline(dev, reps=10):
    """The unmodified reference modules on the same B200 (eval, no_grad): PyTorch eager fp32 and bf16 autocast --
    north_star's '>= 10x the reference single-GPU PyTorch forward' denominator."""
    import torch
    sa, ro = ref_modules(dev)
    feats, init = ref_inputs(dev)
    out = {}
    with torch.no_grad():
        for name, ctx in (('fp32', None), ('bf16_autocast', torch.bfloat16)):
            def run():
                if ctx is None:
                    return ref_step(sa, ro, feats, init)
                with torch.autocast('cuda', dtype=ctx):
                    return ref_step(sa, ro, feats, init)
            for _ in range(3):
                run()
            torch.cuda.synchronize(dev)
            ms = []
            for _ in range(reps):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                run()
                b.record()
                torch.cuda.synchronize(dev)
                ms.append(a.elapsed_time(b))
            out[name] = {'ms_per_step': float(np.median(ms)), 'value': frames_per_step() / (float(np.median(ms)) * 1e-3),
                         'unit': UNIT}
    out['what'] = ('unmodified reference SlotAttention (per time step, as StoSAVi.encode) + SlotRollouter on this GPU, '
                   f'torch {torch.__version__} eager, eval / no_grad, device-resident inputs, median of {reps}')
    del sa, ro, feats, init
    torch.cuda.empty_cache()
    return out


def run_reference(args):
    """--impl reference: the reference's own implementation (oracle/_ref: unmodified savi.py / slotformer.py) on the
    box's host cores; every step is the full 64-clip workload, nothing is extrapolated."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    ts = cpu_reference_times(args.steps, args.warmup)
    t = float(np.mean(ts))
    fps = frames_per_step() / t
    sample = (f'each step = the full workload ({WL["B"]} clips: {WL["B"] * WL["T_in"]} frames of Slot Attention + '
              f'{WL["B"]}-clip rollout) through the unmodified reference modules, torch {_torch_version()} CPU fp32')
    line = {'metric': METRIC, 'value': fps, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': t * 1e3, 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'impl': 'reference',
            'config': {'workload': WORKLOAD, 'ms_min': float(np.min(ts)) * 1e3, 'ms_max': float(np.max(ts)) * 1e3},
            'cpu_baseline': {'value': fps, 'unit': UNIT, 'cores': cpu_threads(), 'kind': 'reference',
                             'sample': sample},
            'e2e': {'value': fps, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'gpu_launches': 0}
    print(json.dumps(line))


def build_savi(dev):
    """StoSAVi of the OBJ3D config (savi_obj3d_params.py:38-72) at 128 x 128, K = 6, random-init weights."""
    import torch
    from slotformer_b200.base_slots.models import StoSAVi
    K, D = WL['K'], WL['D']
    torch.manual_seed(0)
    return StoSAVi(
        resolution=(128, 128), clip_len=WL['T_in'],
        slot_dict=dict(num_slots=K, slot_size=D, slot_mlp_size=WL['Dm'], num_iterations=WL['iters'], kernel_mlp=True),
        enc_dict=dict(enc_channels=(3, 64, 64, 64, 64), enc_ks=5, enc_out_channels=WL['C'], enc_norm=''),
        dec_dict=dict(dec_channels=(D, 64, 64, 64, 64), dec_resolution=(8, 8), dec_ks=5, dec_norm=''),
        pred_dict=dict(pred_type='transformer', pred_rnn=True, pred_norm_first=True, pred_num_layers=2,
                       pred_num_heads=4, pred_ffn_dim=4 * D, pred_sg_every=None),
        loss_dict=dict(use_post_recon_loss=True, kld_method='none')).to(dev).eval()


def model_variants(dev, ro, iters=7):
    """next to the hot-path figure (H, the headline `value`) the same shapes through the model
    classes, device-resident inputs, this GPU only.  (E): images -> StoSAVi.encode (cuDNN CNN, fused encoder tail,
    then per frame the transition kernel + Slot Attention -- the frames of a clip are a serial chain in SAVi, unlike
    (H) where the 384 frames are one batch) -> SlotRollouter.  (E+D): plus StoSAVi.decode of the predicted slots
    (cuDNN deconvolutions + the decode_combine kernel)."""
    import torch
    from slotformer_b200 import engine
    B, T_in, T_out, K, D = (WL[k] for k in ('B', 'T_in', 'T_out', 'K', 'D'))
    savi = build_savi(dev)
    img = torch.rand((B, T_in, 3, 128, 128), device=dev, generator=torch.Generator(device=dev).manual_seed(5)) * 2 - 1

    def run_e(decode):
        savi._reset_rnn()
        _, slots, _ = savi.encode(img)
        pred = ro(slots, T_out)
        if decode:
            flat = pred.flatten(0, 1)
            for i in range(0, flat.shape[0], 128):          # decoder activations: 128 frames x K slots per slice
                savi.decode(flat[i:i + 128])
        return pred

    out = {}
    with torch.no_grad():
        for key, decode in (('E', False), ('E+D', True)):
            for _ in range(2):
                run_e(decode)
            torch.cuda.synchronize(dev)
            n0 = engine.launch_count()
            times = []
            for _ in range(iters):                       # per-call events, median: the host launches ~100 cuDNN / glue kernels per call
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                pred = run_e(decode)
                b.record()
                torch.cuda.synchronize(dev)
                times.append(a.elapsed_time(b))
            assert torch.isfinite(pred).all()
            ms = float(np.median(times))
            out[key] = {'ms_per_step': ms, 'value': frames_per_step() / (ms * 1e-3), 'unit': UNIT,
                        'sfb_launches_per_step': (engine.launch_count() - n0) // iters}
    out['what'] = ('(E) images [64, 6, 3, 128, 128] -> StoSAVi.encode (cuDNN CNN + sfb encoder tail, then 6 serial frames of '
                   'sfb transition + Slot Attention on 64 frames each) -> sfb rollout; (E+D) + StoSAVi.decode of the 640 predicted '
                   'frames (cuDNN deconvolutions + sfb decode_combine); device-resident inputs, random-init weights')
    return out


def config1_variant(dev, iters=20):
    """BASELINE.json configs[0] -- 'OBJ3D SAVi slot extraction, B=4, 64x64, K=5, 3 iters' (the reference's own CPU-runnable
    case): StoSAVi(testing=True) on images [4, 6, 3, 64, 64], device-resident; ours (one CUDA-graph replay of the frame
    loop: per frame the transition kernel + Slot Attention) next to the UNMODIFIED reference StoSAVi (oracle/_ref) with the
    same state_dict on this GPU (PyTorch eager) and on the host cores.  Extracted slots compared in passing."""
    import torch
    from oracle.build_ref import import_ref
    from slotformer_b200.base_slots.models import StoSAVi
    kw = dict(
        resolution=(64, 64), clip_len=6,
        slot_dict=dict(num_slots=5, slot_size=128, slot_mlp_size=256, num_iterations=3, kernel_mlp=True),
        enc_dict=dict(enc_channels=(3, 64, 64, 64, 64), enc_ks=5, enc_out_channels=128, enc_norm=''),
        dec_dict=dict(dec_channels=(128, 64, 64, 64, 64), dec_resolution=(8, 8), dec_ks=5, dec_norm=''),
        pred_dict=dict(pred_type='transformer', pred_rnn=True, pred_norm_first=True, pred_num_layers=2,
                       pred_num_heads=4, pred_ffn_dim=512, pred_sg_every=None),
        loss_dict=dict(use_post_recon_loss=True, kld_method='none'))
    torch.manual_seed(0)
    ours = StoSAVi(**kw).to(dev).eval()
    ref = import_ref()['StoSAVi'](**kw).eval()
    ref.load_state_dict(ours.state_dict(), strict=True)
    ours.testing = ref.testing = True
    img = torch.rand((4, 6, 3, 64, 64), generator=torch.Generator().manual_seed(3)) * 2 - 1
    d_img = img.to(dev)
    frames = 4 * 6

    def gpu_ms(model):
        with torch.no_grad():
            for _ in range(3):
                out = model({'img': d_img})['post_slots']
            torch.cuda.synchronize(dev)
            ts = []
            for _ in range(iters):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                out = model({'img': d_img})['post_slots']
                b.record()
                torch.cuda.synchronize(dev)
                ts.append(a.elapsed_time(b))
        return float(np.median(ts)), out

    t_ours, s_ours = gpu_ms(ours)
    ref = ref.to(dev)
    t_ref_gpu, s_ref = gpu_ms(ref)
    err = float((s_ours - s_ref).abs().max() / s_ref.abs().max())
    ref = ref.cpu()
    torch.set_num_threads(cpu_threads())
    with torch.no_grad():
        ref({'img': img})
        cs = []
        for _ in range(3):
            t0 = time.perf_counter()
            ref({'img': img})
            cs.append(time.perf_counter() - t0)
    t_cpu = float(np.median(cs)) * 1e3
    return {'workload': 'OBJ3D SAVi slot extraction, B=4, 64x64, K=5, 3 iters, T=6 (StoSAVi.forward, testing=True)',
            'ours': {'ms_per_step': t_ours, 'value': frames / (t_ours * 1e-3), 'unit': UNIT},
            'reference_gpu_eager': {'ms_per_step': t_ref_gpu, 'value': frames / (t_ref_gpu * 1e-3), 'unit': UNIT},
            'reference_cpu': {'ms_per_step': t_cpu, 'value': frames / (t_cpu * 1e-3), 'unit': UNIT, 'cores': cpu_threads()},
            'slots_rel_err_vs_reference_gpu': err}


# ---- GPU arm ----------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    from slotformer_b200 import engine
    from slotformer_b200.build import build_extension
    from slotformer_b200.base_slots.models import SlotAttention
    from slotformer_b200.video_prediction.models import SlotRollouter

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm')
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    from slotformer_b200.parallel import bind_to_gpu_numa
    numa_cpus = bind_to_gpu_numa(local)        # before any pinned allocation: host buffers land on the GPU's NUMA node
    # rank 0 prints ONE JSON line on stdout: NCCL's version banner (NCCL_DEBUG=VERSION in this image's environment,
    # printed with printf when the communicator is created) is sent to stderr instead
    saved_stdout = None
    if world > 1:
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        dist.init_process_group('nccl', device_id=dev)
    if rank == 0:
        build_extension()
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize(dev)
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        os.close(saved_stdout)
    engine.load()

    B, T_in, T_out, K, N, C, D = (WL[k] for k in ('B', 'T_in', 'T_out', 'K', 'N', 'C', 'D'))
    sa_w, ro_w = make_weights()
    sa = SlotAttention(C, WL['iters'], K, D, WL['Dm'])
    sa.load_state_dict({k: torch.from_numpy(v) for k, v in sa_w.items()})
    ro = SlotRollouter(K, D, T_in, d_model=WL['d'], num_layers=WL['layers'], num_heads=WL['heads'],
                       ffn_dim=WL['F'])
    ro.load_state_dict({k: torch.from_numpy(v) for k, v in ro_w.items()}, strict=False)
    sa, ro = sa.to(dev).eval(), ro.to(dev).eval()

    gen = torch.Generator(device=dev).manual_seed(1 + rank)
    frames = B * T_in
    feats = torch.randn((frames, N, C), device=dev, generator=gen)
    feats.mul_(0.5 + 1.5 * torch.rand((frames, N, 1), device=dev, generator=gen))
    init = torch.randn((frames, K, D), device=dev, generator=gen)

    def step(f, s0):
        slots = sa(f, s0)                                   # [B*T_in, K, D]
        pred = ro(slots.view(B, T_in, K, D), T_out)         # [B, T_out, K, D]
        return slots, pred

    def sync_all():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    with torch.no_grad():
        for _ in range(max(3, args.warmup)):
            step(feats, init)
        sync_all()

        # ---------------- device-resident timing ----------------
        # (1) serial: one batch at a time on one stream (per-batch latency; kernel durations with every SM)
        sa_ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True),
                  torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        sync_all()
        for i in range(args.steps):
            sa_ev[i][0].record()
            slots = sa(feats, init)
            sa_ev[i][1].record()
            ro(slots.view(B, T_in, K, D), T_out)
            sa_ev[i][2].record()
        torch.cuda.synchronize(dev)
        serial_sa_ms = float(np.mean([a.elapsed_time(b) for a, b, _ in sa_ev]))
        serial_ro_ms = float(np.mean([b.elapsed_time(c) for _, b, c in sa_ev]))
        serial_ms = float(sa_ev[0][0].elapsed_time(sa_ev[-1][2])) / args.steps

        # (2) per-kernel durations inside the pipeline (eager launches on the two streams, timing events; median)
        pipe = engine.HotPathPipeline(sa, ro, dev, clips=B)
        with pipe:
            for _ in range(max(3, args.warmup)):
                pipe.submit(feats, init, B, T_in, T_out)
        sync_all()
        timing = []
        with pipe:
            for i in range(args.steps):
                pipe.submit(feats, init, B, T_in, T_out, timing=timing)
        torch.cuda.synchronize(dev)
        sa_ms = float(np.median([t[0].elapsed_time(t[1]) for t in timing]))
        ro_ms = float(np.median([t[2].elapsed_time(t[3]) for t in timing]))

        # (3) the timed region: K batches through the two-stage pipeline (Slot Attention of batch i+1 overlaps the
        #     rollout of batch i on a second stream; same kernels, same per-batch results), captured as ONE CUDA
        #     graph so that the host issues a single launch and cannot fall behind the GPU
        n0 = engine.launch_count()
        graph, outs = pipe.capture([(feats, init)] * args.steps, B, T_in, T_out)
        launches = engine.launch_count() - n0 - 2 * 6          # minus the two warm-up batches inside capture()
        graph.replay()                                         # warm-up: one replay = K >= W steps
        sync_all()
        sampler = ClockSampler(local)
        sampler.start()
        t_wait = time.perf_counter()
        while not sampler.sm and sampler.err is None and time.perf_counter() - t_wait < 5.0:
            time.sleep(0.001)                                  # NVML is initialised: samples now arrive every ~2 ms
        sampler.sm.clear()
        graph.replay()                                         # load for the clock samples before the timed replays too
        # REPLAYS timed regions of exactly K steps each (one graph replay, barrier + synchronize on both sides, max
        # over ranks); the line reports the median region, min / max beside it
        REPLAYS = 5
        regions = []
        for _ in range(REPLAYS):
            sync_all()
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
            ev[0].record()
            graph.replay()                                     # exactly K steps
            ev[1].record()
            torch.cuda.synchronize(dev)
            regions.append(ev[0].elapsed_time(ev[1]))
        clocks = sampler.result()
        assert torch.isfinite(outs[-1][1]).all()
        t = torch.tensor(regions, device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        regions = [float(x) for x in t.tolist()]
        ms_total = float(np.median(regions))
        ms_per_step = ms_total / args.steps
        value = world * frames_per_step() / (ms_per_step * 1e-3)

        # ---------------- end-to-end with host buffers ----------------
        # Same pipeline fed from pinned HOST buffers: every step copies its feature grids + initial slots
        # host->device (8 chunks, Slot Attention starts on a chunk as soon as it has landed) and reads the
        # extracted + predicted slots back; two steps in flight (double-buffered device staging).
        e2e = None
        if not args.no_e2e:
            tail = engine.EncoderTailEngine()
            tail_w = {k: torch.from_numpy(v).to(dev) for k, v in make_tail_weights().items()}
            copy_stream = torch.cuda.Stream(dev)
            out_stream = torch.cuda.Stream(dev)
            nchunk = 8
            cf = frames // nchunk
            h_init = torch.empty((frames, K, D), dtype=torch.float32, pin_memory=True)
            h_init.copy_(init)
            h_slots = [torch.empty((frames, K, D), dtype=torch.float32, pin_memory=True) for _ in range(2)]
            h_pred = [torch.empty((B, T_out, K, D), dtype=torch.float32, pin_memory=True) for _ in range(2)]
            d_init = [torch.empty_like(init) for _ in range(2)]
            d_slots = [torch.empty_like(init) for _ in range(2)]

            def e2e_route(route):
                """route 'cnn': host CNN output -> encoder tail -> Slot Attention on operand tiles -> rollout;
                route 'grid': host fp32 feature grid -> Slot Attention -> rollout (round 1)."""
                if route == 'cnn':
                    h_in = torch.empty((frames, 64, 64, 64), dtype=torch.float32, pin_memory=True)
                    h_in.copy_(torch.randn(h_in.shape, generator=torch.Generator().manual_seed(7 + rank)))
                else:
                    h_in = torch.empty((frames, N, C), dtype=torch.float32, pin_memory=True)
                    h_in.copy_(feats)
                d_in = [torch.empty(h_in.shape, dtype=torch.float32, device=dev) for _ in range(2)]
                done_ev = [None, None]

                def submit(i):
                    sl = i & 1
                    if done_ev[sl] is not None:
                        done_ev[sl].synchronize()            # staging buffers of step i-2 are free again
                    evs = []
                    with torch.cuda.stream(copy_stream):
                        d_init[sl].copy_(h_init, non_blocking=True)
                        for c in range(nchunk):
                            d_in[sl][c * cf:(c + 1) * cf].copy_(h_in[c * cf:(c + 1) * cf], non_blocking=True)
                            e = torch.cuda.Event()
                            e.record(copy_stream)
                            evs.append(e)
                    with torch.cuda.stream(pipe.s_sa):
                        for c in range(nchunk):
                            pipe.s_sa.wait_event(evs[c])
                            x = d_in[sl][c * cf:(c + 1) * cf]
                            if route == 'cnn':
                                x = tail.forward(x, tail_w, C, max_ctas=pipe.sa_ctas)
                            d_slots[sl][c * cf:(c + 1) * cf] = sa(x, d_init[sl][c * cf:(c + 1) * cf])
                        ready = torch.cuda.Event()
                        ready.record(pipe.s_sa)
                    pipe.s_ro.wait_event(ready)
                    with torch.cuda.stream(pipe.s_ro):
                        pred = ro(d_slots[sl].view(B, T_in, K, D), T_out)
                        rdone = torch.cuda.Event()
                        rdone.record(pipe.s_ro)
                    pred.record_stream(out_stream)
                    out_stream.wait_event(rdone)
                    with torch.cuda.stream(out_stream):
                        h_slots[sl].copy_(d_slots[sl], non_blocking=True)
                        h_pred[sl].copy_(pred, non_blocking=True)
                        fin = torch.cuda.Event()
                        fin.record(out_stream)
                    done_ev[sl] = fin

                def run(n):
                    with pipe:
                        for i in range(n):
                            submit(i)
                        for e in done_ev:
                            if e is not None:
                                e.synchronize()
                    torch.cuda.synchronize(dev)

                n0 = engine.launch_count()
                run(2)
                per_step = (engine.launch_count() - n0) // 2
                sync_all()
                done_ev[0] = done_ev[1] = None
                steps_ = max(3, min(args.steps, 10))
                t0 = time.perf_counter()
                run(steps_)
                dt = (time.perf_counter() - t0) / steps_
                tt = torch.tensor([dt], device=dev, dtype=torch.float64)
                if world > 1:
                    dist.all_reduce(tt, op=dist.ReduceOp.MAX)
                dt = float(tt.item())
                assert torch.isfinite(h_pred[0]).all() and torch.isfinite(h_slots[0]).all()
                res = {'value': world * frames_per_step() / dt, 'unit': UNIT,
                       'h2d_bytes_per_step': int(h_in.numel() * 4 + h_init.numel() * 4),
                       'd2h_bytes_per_step': int(h_slots[0].numel() * 4 + h_pred[0].numel() * 4),
                       'ms_per_step': dt * 1e3, 'steps': steps_, 'gpu_launches_per_step': int(per_step)}
                del h_in, d_in
                torch.cuda.empty_cache()
                return res

            e2e = e2e_route('cnn')
            e2e['note'] = (f'pinned host buffers; host input = CNN encoder output [{frames}, 64, 64, 64] fp32 (savi.py:367), '
                           f'{nchunk}-chunk H2D overlapped with encoder tail + Slot Attention on operand tiles, rollout, D2H of '
                           'slots + predictions; two steps in flight (PCIe-bound)')
            grid = e2e_route('grid')
            grid['note'] = 'round-1 route: host fp32 feature grid [384, 4096, 128] (806 MB per step), no encoder tail on the device'
            e2e['feature_grid_route'] = grid
            # The user-level call one step further up: host IMAGES -> StoSAVi.encode (cuDNN CNN, sfb encoder tail, per
            # frame sfb transition + Slot Attention) -> sfb rollout -> host.  5.3x fewer bytes cross PCIe, the device does
            # the CNN: the route that is not bound by the host's aggregate H2D bandwidth when 8 ranks share it.
            savi = build_savi(dev)
            h_img = torch.empty((B, T_in, 3, 128, 128), dtype=torch.float32, pin_memory=True)
            h_img.copy_(torch.rand(h_img.shape, generator=torch.Generator().manual_seed(11 + rank)) * 2 - 1)
            h_s = torch.empty((B, T_in, K, D), dtype=torch.float32, pin_memory=True)
            h_p = torch.empty((B, T_out, K, D), dtype=torch.float32, pin_memory=True)
            from slotformer_b200.pipeline import ClipPipeline

            def image_steps(n, clip_pipe):
                for _ in range(n):
                    clip_pipe.submit(h_img, h_s, h_p)
                clip_pipe.drain()

            # one eager step to count the library's kernels (graph replays hide them), then the pipeline of the timed run
            savi.use_cuda_graph = False
            counting = ClipPipeline(savi, ro, T_out, dev, capture_encode=False)
            counting.warmup(h_img)
            n0 = engine.launch_count()
            image_steps(1, counting)
            img_launches = engine.launch_count() - n0
            del counting
            savi.use_cuda_graph = True
            clip_pipe = ClipPipeline(savi, ro, T_out, dev)
            clip_pipe.warmup(h_img)
            image_steps(3, clip_pipe)
            sync_all()
            n_img = max(3, min(args.steps, 10))
            t0 = time.perf_counter()
            image_steps(n_img, clip_pipe)
            dt = (time.perf_counter() - t0) / n_img
            tt = torch.tensor([dt], device=dev, dtype=torch.float64)
            if world > 1:
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            dt = float(tt.item())
            assert torch.isfinite(h_p).all() and torch.isfinite(h_s).all()
            image = {
                'value': world * frames_per_step() / dt, 'unit': UNIT, 'h2d_bytes_per_step': int(h_img.numel() * 4),
                'd2h_bytes_per_step': int(h_s.numel() * 4 + h_p.numel() * 4), 'ms_per_step': dt * 1e3, 'steps': n_img,
                'gpu_launches_per_step': int(img_launches),
                'note': 'the user-level call: pinned host images [64, 6, 3, 128, 128] fp32 -> StoSAVi.encode (cuDNN CNN + sfb encoder '
                        'tail, 6 serial frames of sfb transition + Slot Attention, one CUDA-graph replay) -> sfb rollout -> host slots + '
                        'predictions; slotformer_b200.pipeline.ClipPipeline: three streams, H2D and CNN + tail (one CUDA graph) of step i+1 overlap the '
                        'frame loop + rollout of step i'}
            # headline e2e = the image route; the two routes that enter at the operator boundary (PCIe-bound: 404 / 806 MB
            # per step, and bound by the host's aggregate H2D bandwidth when 8 ranks share it) are kept beside it
            image['cnn_output_route'] = {k: v for k, v in e2e.items() if k != 'feature_grid_route'}
            image['feature_grid_route'] = e2e['feature_grid_route']
            e2e = image
            del savi, clip_pipe
            torch.cuda.empty_cache()
            e2e['host_affinity'] = (f'{len(numa_cpus)} CPUs local to GPU {local} (NVML), set before the pinned allocations'
                                    if numa_cpus else 'not bound')

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    pk = peaks()
    sa_gbs = frames * sa_bytes_per_frame() / (sa_ms * 1e-3) / 1e9
    sa_gbs_serial = frames * sa_bytes_per_frame() / (serial_sa_ms * 1e-3) / 1e9
    ro_tf = ro_flops_total() / (ro_ms * 1e-3) / 1e12
    line = {
        'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
        'warmup': max(3, args.warmup), 'ms_per_step': ms_per_step, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f16', 'data': 'synthetic',
        'config': {'workload': WORKLOAD,
                   'arithmetic': 'fp16 tensor-core operands, fp32 accumulate / LayerNorm / softmax / GRU',
                   'l2': 'no flush needed: each step streams 805 MB of features (> 126 MB L2)',
                   'per_gpu_clips': B, 'parallelism': f'clip-sharded x{world}, no data-path collective',
                   'pipeline': 'two-stage over consecutive batches: Slot Attention of batch i+1 (84 SMs) runs '
                               'concurrently with the rollout of batch i (one SM per clip); per-batch results unchanged; '
                               'the K timed steps are one CUDA-graph replay (two streams captured)',
                   'serial_ms_per_step': serial_ms, 'serial_sa_ms': serial_sa_ms, 'serial_rollout_ms': serial_ro_ms},
        'clocks': clocks, 'gpu_launches': int(launches),
        'roofline': {'kernel': 'sfb_sa_forward: sa_update x3 + sa_pass_tc_first + sa_pass_tc_next (tcgen05 passes)',
                     'bound': 'hbm', 'achieved': sa_gbs, 'peak': pk['hbm'],
                     'unit': 'GB/s', 'frac': sa_gbs / pk['hbm'], 'traffic': SA_TRAFFIC_BYTES,
                     'traffic_source': SA_TRAFFIC_SOURCE,
                     'ms_per_launch': sa_ms, 'peak_source': pk['src'],
                     'algorithmic_bytes_per_launch': frames * sa_bytes_per_frame(),
                     'note': 'timed region = batch pipeline: Slot Attention runs on 84 of 148 SMs next to the rollout',
                     'alone_on_all_sms': {'ms_per_launch': serial_sa_ms, 'achieved': sa_gbs_serial,
                                          'frac': sa_gbs_serial / pk['hbm']}},
        'roofline_rollout': {'kernel': 'ro_umma_forward_kernel (tcgen05)', 'bound': 'tensor', 'achieved': ro_tf,
                             'peak': pk['tf'], 'unit': 'TFLOP/s', 'frac': ro_tf / pk['tf'],
                             'ms_per_launch': ro_ms, 'flops_per_launch': ro_flops_total()},
    }
    line['timed_regions'] = {'n': len(regions), 'steps_each': args.steps, 'ms_per_step_min': min(regions) / args.steps,
                             'ms_per_step_median': ms_per_step, 'ms_per_step_max': max(regions) / args.steps}
    if e2e is not None:
        line['e2e'] = e2e
    if world == 1 and not args.no_cpu_baseline:
        if not args.no_variants:
            line['variants'] = model_variants(dev, ro)
            line['variants']['config1'] = config1_variant(dev)
        line['gpu_eager_baseline'] = gpu_eager_baseline(dev)
        line['cpu_baseline'] = cpu_baseline()
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=50)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-variants', action='store_true')
    args = ap.parse_args()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()
