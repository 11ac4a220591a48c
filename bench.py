#!/usr/bin/env python
"""Benchmark of the DMC-Net train step (BASELINE.json metric: clips/sec,
224x224, 3 segments, DMC generator + ResNet-18, train step).

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--config dmcnet|gan]
                    [--scaling weak|strong] [--no-extras]

N>1 is launched by the driver through torch.distributed.run (one rank per GPU,
NCCL); rank 0 prints ONE JSON line.

Headline (`value`, `e2e`, `roofline`, `cpu_baseline`): BASELINE config 2 -- "dmcnet (no GAN) train
step, synthetic HMDB-51-shaped batch=64 per GPU, 3 segments, flow-MSE + CE loss" -- weak scaling
(64 clips per rank, one sum all-reduce of the gradient bucket per step).

The same line carries, under `configs`, the other single-node configurations BASELINE.json names,
measured in the same process right after the headline with the same W / K:

  configs.config3        dmcnet_GAN step pair (D-step + G-step, Discriminator3, 101 classes), B=64 per GPU
  configs.config4_strong dmcnet_GAN, 51 classes, GLOBAL batch 512 fixed and sharded 512/N per GPU
                         (strong scaling; one GPU holds all 512 clips at N=1)
  configs.config5_i3d    dmcnet_I3D (DenseNetTiny estimator + I3D), 16-frame clips, GLOBAL batch 32 sharded 32/N
                         per GPU (the reference's nn.DataParallel split of --batch-size 32), SGD-Nesterov

  value      clips/s with the batch already resident in HBM (CUDA events, max over ranks)
  e2e        clips/s through FusedTrainStep.step_pipelined_u8() with HOST (pinned) input: the uint8
             sample stack [B,S,H,W,7] (the format CoviarDataSet's augmentation produces) crosses PCIe
             every step, is split / normalised on the device, and the metric record is read back
  roofline   the dominant kernel family, timed live with CUDA events on the launch stream
  cpu_baseline / --impl reference
             the reference's OWN Model (oracle/_ref, byte-compiled from /root/reference by
             oracle/make_ref.py) inside the restated step of its train.py (oracle/ref_step.py), torch
             CPU fp32, all host threads: kind "reference"; if oracle/_ref is absent, the oracle port
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CONFIGS = {
    'dmcnet': dict(gan=False, arch_d=None, num_class=51,
                   workload='dmcnet train step (BASELINE config 2): DenseNetTiny generator + ResNet-18, '
                            'flow-MSE + CE, Adam, B=%d clips x 3 segments x 224x224 per GPU, 51 classes'),
    'gan': dict(gan=True, arch_d='Discriminator3', num_class=101,
                workload='dmcnet_GAN train step (BASELINE config 3): mean of one D-step and one G-step, '
                         'Discriminator3, B=%d clips x 3 segments x 224x224 per GPU, 101 classes'),
    'gan51': dict(gan=True, arch_d='Discriminator3', num_class=51,
                  workload='dmcnet_GAN data-parallel (BASELINE config 4): D-step + G-step pair, Discriminator3, '
                           'HMDB-51 shape (51 classes), B=%d clips x 3 segments x 224x224 per GPU'),
}

# algorithmic FLOPs per clip (dead work removed): 1 MAC = 2 FLOP
GFLOP_PER_CLIP = {'dmcnet': 35.2, 'gan': 35.6, 'gan51': 35.6}
STRONG_GLOBAL_BATCH = 512


def dist_env():
    return int(os.environ.get('RANK', 0)), int(os.environ.get('LOCAL_RANK', 0)), \
        int(os.environ.get('WORLD_SIZE', 1))


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag = index, [], set(), False
        self.max_mhz = None

    def run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            names = {
                getattr(nv, 'nvmlClocksThrottleReasonHwSlowdown', 0x8): 'hw_slowdown',
                getattr(nv, 'nvmlClocksThrottleReasonHwThermalSlowdown', 0x40): 'hw_thermal_slowdown',
                getattr(nv, 'nvmlClocksThrottleReasonSwThermalSlowdown', 0x20): 'sw_thermal_slowdown',
                getattr(nv, 'nvmlClocksThrottleReasonSwPowerCap', 0x4): 'sw_power_cap',
            }
            while not self.stop_flag:
                self.samples.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
                time.sleep(0.05)
        except Exception as e:  # noqa: BLE001
            self.reasons.add('nvml_unavailable:%s' % type(e).__name__)

    def summary(self):
        s = sorted(self.samples)
        return {'sm_mhz': (s[len(s) // 2] if s else None), 'sm_max_mhz': self.max_mhz,
                'reasons': sorted(self.reasons)}


# ---------------------------------------------------------------------------- CPU reference arm
def cpu_sample_batch(requested):
    """B=64 when the host has the memory for it (saved activations ~20 MB/frame), else 16."""
    if requested:
        return requested
    try:
        import psutil
        return 64 if psutil.virtual_memory().available > 48 * (1 << 30) else 16
    except Exception:  # noqa: BLE001
        return 16


def cpu_reference_rate(cfg_name, steps, warmup, sample_batch):
    """The reference's CPU implementation of the step on all host threads: `warmup` untimed + `steps`
    timed iterations (GAN: D+G pairs) of B=sample_batch clips; returns the MEDIAN step rate."""
    import torch
    from oracle import dmc_oracle as O
    from oracle import ref_loader as R
    cfg = CONFIGS[cfg_name]
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sd = O.build_state(cfg['num_class'], cfg['arch_d'], seed=1)
    if R.reference_available():
        from oracle.ref_step import ReferenceTrainer
        tr, kind = ReferenceTrainer(cfg['num_class'], O.HParams(), gan=cfg['gan'], arch_d=cfg['arch_d'], state=sd), \
            'reference'
    else:
        tr, kind = O.OracleTrainer(sd, O.HParams(), gan=cfg['gan'], arch_d=cfg['arch_d']), 'port'
    flow, mv, res, target = O.make_inputs(sample_batch, 3, cfg['num_class'], seed=0)
    per = 2 if cfg['gan'] else 1
    for _ in range(warmup * per):
        tr.step(flow, mv, res, target)
    times = []
    for _ in range(steps):
        t0 = time.perf_counter()
        for _ in range(per):
            tr.step(flow, mv, res, target)
        times.append((time.perf_counter() - t0) / per)
    times.sort()
    dt = times[len(times) // 2]
    sample = ('median of %d timed steps after %d warm-up, B=%d clips x 3 segments of the same workload, '
              'torch CPU fp32, %d threads' % (steps, warmup, sample_batch, cores))
    return {'value': sample_batch / dt, 'unit': 'clips/s', 'cores': cores, 'kind': kind, 'sample': sample,
            'ms_per_step': dt * 1e3, 'batch': sample_batch}


def run_reference(args):
    rank, _, world = dist_env()
    if rank != 0:
        return
    cfg = CONFIGS[args.config]
    sb = cpu_sample_batch(args.cpu_sample_batch)
    steps, warmup = max(5, min(args.steps, 6)), max(1, min(args.warmup, 1))
    base = cpu_reference_rate(args.config, steps, warmup, sb)
    rate = base['value']
    line = {
        'impl': 'reference', 'metric': 'clips/sec', 'value': rate, 'unit': 'clips/s', 'n_gpus': args.gpus,
        'steps': steps, 'warmup': warmup, 'ms_per_step': base['ms_per_step'], 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': cfg['workload'] % sb, 'clips_per_gpu': sb, 'global_batch': sb, 'segments': 3,
                   'note': 'CPU, torch fp32: the reference Model (oracle/_ref) inside the restated train.py step'
                   if base['kind'] == 'reference' else 'CPU, torch fp32, oracle restatement of the reference step'},
        'cpu_baseline': {k: base[k] for k in ('value', 'unit', 'cores', 'kind', 'sample')},
        'e2e': {'value': rate, 'unit': 'clips/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    emit(line)


def emit(line: dict):
    """The ONE JSON line goes to the real stdout; everything else printed by libraries
    (NCCL banner, torchrun notices) was redirected to stderr by protect_stdout()."""
    os.write(_REAL_STDOUT, (json.dumps(line) + '\n').encode())


_REAL_STDOUT = 1


def protect_stdout():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)


# ---------------------------------------------------------------------------- one measured configuration
def measure(cfg_name, B, args, *, rank, local_rank, world, clocks=False, roofline=False):
    """W warm-up + K timed steps of one configuration at B clips per rank: resident throughput,
    end-to-end throughput from pinned host memory, optional per-family breakdown."""
    import torch
    import torch.distributed as dist
    from dmcnet_b200 import ops
    from dmcnet_b200.engine import DmcEngine
    from dmcnet_b200.model import build_state
    from dmcnet_b200.trainer import FusedTrainStep, HParams

    cfg = CONFIGS[cfg_name]
    S, H, W = 3, 224, 224
    per = 2 if cfg['gan'] else 1                      # GAN: one "step" = D-step + G-step pair / 2

    # synthetic inputs with the CoviarDataSet value model (uint8 -> normalised fp32), seeded per rank.
    # Drawn once as the uint8 stack [B,S,H,W,7] (flow | mv | residual), normalised as dataset.py:251-263.
    g = torch.Generator().manual_seed(1234 + rank)

    def u8(ch, sigma):
        return torch.clamp(torch.round(128.0 + sigma * torch.randn((B, S, H, W, ch), generator=g)), 0, 255)
    stack_u8 = torch.cat((u8(2, 30.0), u8(2, 25.0), u8(3, 20.0)), 4).to(torch.uint8).contiguous().pin_memory()
    target = torch.randint(0, cfg['num_class'], (B,), generator=g).pin_memory()

    sd = build_state(cfg['num_class'], cfg['arch_d'], seed=1)
    eng = DmcEngine(cfg['num_class'], S, B * S, gan=cfg['gan'], arch_d=cfg['arch_d'])
    eng.load_state(sd)
    tr = FusedTrainStep(eng, HParams(), B, world_size=world, use_graph=not args.no_graph, pipelined=True,
                        graph_allreduce=args.graph_allreduce, overlap=(False if args.no_overlap else None))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], device='cuda')
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0])

    # ---------------- device-resident throughput
    tr.load_inputs_u8(stack_u8, target)               # split + normalise on the device, once
    masks_d = eng.draw_dropout_masks(2 * B * S) if cfg['gan'] else None
    masks_g = eng.draw_dropout_masks(B * S) if cfg['gan'] else None

    def resident_step():
        mode = tr._mode()
        if cfg['gan']:
            m = 2 * eng.N if mode == 'D' else eng.N
            eng.set_masks(masks_d if mode == 'D' else masks_g, m)
        tr._run(mode, True)
        tr.iteration += 1

    for _ in range(args.warmup * per + 2 * per):      # +2: graph warm-up and capture
        resident_step()
    ops.reset_launch_count()
    sampler = None
    if clocks:
        sampler = ClockSampler(local_rank)
        sampler.start()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps * per):
        resident_step()
    e1.record()
    barrier()
    ms_step = max_over_ranks(e0.elapsed_time(e1) / args.steps / per)
    launches = ops.launch_count() if not tr.use_graph else tr.launches_per_step * args.steps * per

    # ---------------- end to end through the public API with host inputs
    # every step copies ITS batch from pinned host memory (copy stream, overlapping the previous
    # step's compute), runs the input kernels + the step, and reads back the metric record of the
    # step that just finished; flush() inside the timed region collects the last one.
    def e2e_step():
        mk = masks_d if tr._mode() == 'D' else masks_g
        return tr.step_pipelined_u8(stack_u8, target, masks=mk)
    for _ in range(2 * per):
        e2e_step()
    tr.flush()
    barrier()
    e0.record()
    last = None
    for _ in range(args.steps * per):
        m = e2e_step()
        last = m or last
    last = tr.flush() or last
    e1.record()
    barrier()
    ms_e2e = max_over_ranks(e0.elapsed_time(e1) / args.steps / per)
    if sampler is not None:
        sampler.stop_flag = True
        sampler.join(timeout=2)
    h2d = stack_u8.numel() + target.numel() * 8
    d2h = 16 * 8                                     # one pinned 16-double stats record per step

    roof = None
    if roofline:
        # every rank runs it (the steps contain the gradient all-reduce); rank 0 reports
        from dmcnet_b200.profiling import measure_roofline
        roof = measure_roofline(resident_step, per, min(args.steps, 3), tr)
        barrier()

    clips = B * world
    out = {
        'workload': cfg['workload'] % B, 'clips_per_gpu': B, 'global_batch': clips,
        'value': clips / (ms_step * 1e-3), 'unit': 'clips/s', 'ms_per_step': ms_step,
        'e2e': {'value': clips / (ms_e2e * 1e-3), 'unit': 'clips/s', 'h2d_bytes_per_step': h2d,
                'd2h_bytes_per_step': d2h, 'ms_per_step': ms_e2e},
        'gpu_launches': int(launches),
        'algorithmic_tflops': GFLOP_PER_CLIP[cfg_name] * clips / ms_step,
        'last_metrics': last, 'cuda_graph': bool(tr.use_graph), 'two_streams': bool(tr.overlap),
    }
    if sampler is not None:
        out['clocks'] = sampler.summary()
    if roof:
        out['roofline'] = roof['roofline']
        out['kernel_breakdown_ms_per_step'] = roof['breakdown']
    del tr, eng, stack_u8
    import gc
    gc.collect()
    torch.cuda.empty_cache()
    return out


# ---------------------------------------------------------------------------- BASELINE config 5 (I3D)
I3D_GLOBAL_BATCH, I3D_CLIP_LEN = 32, 16
# forward 7.31 (estimator, 16 frames) + 51.17 (I3D) GFLOP per clip; a train step is
# counted as 3x the forward (data + weight gradients; the input layers' data gradients are small)
I3D_GFLOP_PER_CLIP = 3.0 * (7.31 + 51.17)


def measure_i3d(B, args, *, rank, local_rank, world):
    """dmcnet_I3D train step (DenseNetTiny estimator + I3D, CE + MSE, SGD-Nesterov, fit()'s non-adversarial
    iteration) at B clips of 16 frames per rank: resident and end-to-end (pinned fp32 sample tensor) rates."""
    import torch
    import torch.distributed as dist
    from dmcnet_b200 import ops
    from dmcnet_b200.i3d_engine import I3DEngine
    from dmcnet_b200.i3d_trainer import I3DHParams, I3DTrainStep
    from dmcnet_b200.i3d_model import build_i3d_state
    T, H, W = I3D_CLIP_LEN, 224, 224
    g = torch.Generator().manual_seed(4321 + rank)
    data = torch.empty(B, 7, T, H, W).normal_(generator=g).pin_memory()
    target = torch.randint(0, 51, (B,), generator=g).pin_memory()
    eng = I3DEngine(51, B, T)
    eng.load_state(build_i3d_state(51, 'DenseNetTiny', seed=1))
    tr = I3DTrainStep(eng, I3DHParams(epoch_thre=0), world_size=world, use_graph=not args.no_graph)
    mask = eng.draw_dropout_mask(0.5, g)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], device='cuda')
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0])
    d_data, d_target = data.cuda(), target.cuda()
    steps = max(2, min(args.steps, 5))
    out = {}
    for key, dd, tt in (('resident', d_data, d_target), ('e2e', data, target)):
        # e2e: every step copies ITS batch from pinned host memory (copy stream, overlapping the previous
        # step's compute) and reads back the metrics of the step that just finished; flush() inside the
        # timed region collects the last one
        run = (lambda: tr.step(dd, tt, dropout_mask=mask, metrics=False)) if key == 'resident' else \
            (lambda: tr.step_pipelined(dd, tt, dropout_mask=mask))
        for _ in range(2):
            run()
        if key == 'e2e':
            tr.flush()
        ops.reset_launch_count()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            last = run()
        if key == 'e2e':
            last = tr.flush()
        e1.record()
        barrier()
        out[key] = max_over_ranks(e0.elapsed_time(e1) / steps)
        if key == 'resident':
            launches = ops.launch_count()
    if tr.use_graph:            # replayed launches do not pass the C ABI's counter: count one eager step
        tr.use_graph = False
        ops.reset_launch_count()
        tr.step(d_data, d_target, dropout_mask=mask, metrics=False)
        torch.cuda.synchronize()
        launches = ops.launch_count() * steps
        tr.use_graph = True
    clips = B * world
    res = {'workload': 'dmcnet_I3D train step (BASELINE config 5): DenseNetTiny estimator + I3D (Inception-3D), CE + MSE, '
                       'SGD-Nesterov, B=%d clips x 16 frames x 224x224 per GPU, 51 classes' % B,
           'clips_per_gpu': B, 'global_batch': clips, 'value': clips / (out['resident'] * 1e-3), 'unit': 'clips/s',
           'ms_per_step': out['resident'], 'steps': steps,
           'e2e': {'value': clips / (out['e2e'] * 1e-3), 'unit': 'clips/s', 'ms_per_step': out['e2e'],
                   'h2d_bytes_per_step': data.numel() * 4 + target.numel() * 8, 'd2h_bytes_per_step': 4 * 4 + 8},
           'gpu_launches': int(launches), 'algorithmic_tflops': I3D_GFLOP_PER_CLIP * clips / out['resident'],
           'last_metrics': last, 'cuda_graph': bool(tr.use_graph)}
    del tr, eng, d_data
    import gc
    gc.collect()
    torch.cuda.empty_cache()
    return res


def cpu_i3d_rate(steps, warmup, sample_batch):
    """The oracle's restatement of the I3D iteration (pinned bit-exactly on the reference's i3d.py) on all host
    threads; median step rate."""
    import torch
    from oracle import i3d_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    tr = O.I3DOracleTrainer(O.build_state(51, 'DenseNetTiny', seed=1), O.I3DHParams(epoch_thre=0))
    data, target = O.make_inputs(sample_batch, I3D_CLIP_LEN, 51, seed=0)
    for _ in range(warmup):
        tr.step(data, target)
    times = []
    for _ in range(steps):
        t0 = time.perf_counter()
        tr.step(data, target)
        times.append(time.perf_counter() - t0)
    times.sort()
    dt = times[len(times) // 2]
    return {'value': sample_batch / dt, 'unit': 'clips/s', 'cores': cores, 'kind': 'port',
            'sample': 'median of %d timed steps after %d warm-up, B=%d clips x 16 frames of the same workload, torch CPU '
                      'fp32, %d threads' % (steps, warmup, sample_batch, cores)}


def main():
    protect_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--config', default='dmcnet', choices=list(CONFIGS), help='headline configuration')
    ap.add_argument('--batch', type=int, default=64, help='clips per GPU (weak) ')
    ap.add_argument('--scaling', default='weak', choices=['weak', 'strong'],
                    help="headline scaling mode; 'strong' fixes the GLOBAL batch at 512 (512/N per GPU)")
    ap.add_argument('--cpu-sample-batch', type=int, default=0, help='0 = 64 if host memory allows, else 16')
    ap.add_argument('--no-graph', action='store_true')
    ap.add_argument('--graph-allreduce', action='store_true',
                    help='capture the NCCL gradient all-reduce inside the step graph (N > 1)')
    ap.add_argument('--no-overlap', action='store_true',
                    help='one stream (default: generator backward / discriminator next to the classifier on two)')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-extras', action='store_true', help='headline only (skip configs.config3 / config4_strong)')
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == 'b200' else args.warmup
    if args.impl == 'reference':
        return run_reference(args)

    import torch
    import torch.distributed as dist

    rank, local_rank, world = dist_env()
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        import datetime
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank),
                                timeout=datetime.timedelta(seconds=300))
    kw = dict(rank=rank, local_rank=local_rank, world=world)
    B = args.batch
    if args.scaling == 'strong':
        if STRONG_GLOBAL_BATCH % world:
            raise SystemExit('strong scaling needs a world size dividing %d' % STRONG_GLOBAL_BATCH)
        B = STRONG_GLOBAL_BATCH // world
    head = measure(args.config, B, args, clocks=True, roofline=True, **kw)

    extras = {}
    if not args.no_extras:
        plan = []
        if args.config != 'gan':
            plan.append(('config3', 'gan', args.batch, 'weak'))
        if STRONG_GLOBAL_BATCH % world == 0:
            plan.append(('config4_strong', 'gan51', STRONG_GLOBAL_BATCH // world, 'strong'))
        if I3D_GLOBAL_BATCH % world == 0:
            plan.append(('config5_i3d', 'i3d', I3D_GLOBAL_BATCH // world, 'strong'))
        for key, cfg_name, b, mode in plan:
            try:
                if cfg_name == 'i3d':
                    r = measure_i3d(b, args, **kw)
                else:
                    r = measure(cfg_name, b, args, roofline=(key == 'config3'), **kw)
                r['scaling'] = mode
                extras[key] = r
            except Exception as e:  # noqa: BLE001  (an extra must never take the headline down)
                extras[key] = {'error': '%s: %s' % (type(e).__name__, str(e)[:300])}
                torch.cuda.empty_cache()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    line = {
        'metric': 'clips/sec', 'value': head['value'], 'unit': 'clips/s', 'n_gpus': world,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': head['ms_per_step'], 'higher_is_better': True,
        'scaling': args.scaling, 'vs_baseline': None,
        'dtype': 'f32 (bf16x3 split on tensor cores, fp32 accumulate)', 'data': 'synthetic',
        'config': {'workload': head['workload'], 'clips_per_gpu': head['clips_per_gpu'],
                   'global_batch': head['global_batch'], 'segments': 3, 'parallelism': 'dp%d' % world,
                   'l2_policy': 'inputs+activations per step (>2 GB) far exceed the 126 MB L2',
                   'cuda_graph': head['cuda_graph'], 'e2e_input': 'uint8 sample stack [B,S,H,W,7], pinned'},
        'e2e': head['e2e'], 'gpu_launches': head['gpu_launches'], 'clocks': head.get('clocks'),
        'algorithmic_tflops': head['algorithmic_tflops'], 'last_metrics': head['last_metrics'],
    }
    if 'roofline' in head:
        line['roofline'] = head['roofline']
        line['kernel_breakdown_ms_per_step'] = head['kernel_breakdown_ms_per_step']
    if extras:
        line['configs'] = extras
    if not args.no_cpu_baseline and world == 1:
        sb = cpu_sample_batch(args.cpu_sample_batch)
        base = cpu_reference_rate(args.config, 5, 1, sb)
        line['cpu_baseline'] = {k: base[k] for k in ('value', 'unit', 'cores', 'kind', 'sample')}
        if 'config3' in extras and 'error' not in extras['config3']:
            b3 = cpu_reference_rate('gan', 5, 1, 16)
            extras['config3']['cpu_baseline'] = {k: b3[k] for k in ('value', 'unit', 'cores', 'kind', 'sample')}
        if 'config5_i3d' in extras and 'error' not in extras['config5_i3d']:
            extras['config5_i3d']['cpu_baseline'] = cpu_i3d_rate(3, 1, 2)
    emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
