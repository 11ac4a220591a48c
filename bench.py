#!/usr/bin/env python
"""Benchmark of the DMC-Net train step (BASELINE.json metric: clips/sec,
224x224, 3 segments, DMC generator + ResNet-18, train step).

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--config dmcnet|gan]

N>1 is launched by the driver through torch.distributed.run (one rank per GPU,
NCCL); rank 0 prints ONE JSON line.  Workload at every N: BASELINE config 2
("dmcnet (no GAN) train step, synthetic HMDB-51-shaped batch=64 per GPU, 3
segments, flow-MSE + CE loss"), weak scaling (64 clips per rank, one sum
all-reduce of the gradient bucket per step).

  value      clips/s with the batch already resident in HBM (CUDA events, max over ranks)
  e2e        clips/s through FusedTrainStep.step() with HOST (pinned) input tensors:
             H2D copy of flow/mv/residual/target and D2H read of the metrics inside
             the timed region
  roofline   the dominant kernel family, timed live with CUDA events on the launch stream
  cpu_baseline / --impl reference
             the CPU restatement of the reference step (oracle/, torch CPU fp32, all
             host threads) on a bounded sample of the same workload
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
                            'flow-MSE + CE, Adam, B=64 clips x 3 segments x 224x224 per GPU, 51 classes'),
    'gan': dict(gan=True, arch_d='Discriminator3', num_class=101,
                workload='dmcnet_GAN train step (BASELINE config 3): mean of one D-step and one G-step, '
                         'Discriminator3, B=64 clips x 3 segments x 224x224 per GPU, 101 classes'),
}

# algorithmic FLOPs per frame: 1 MAC = 2 FLOP
GFLOP_PER_CLIP = {'dmcnet': 35.2, 'gan': 35.6}


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


def cpu_reference_rate(cfg_name, steps, warmup, sample_batch):
    """Reference CPU path: oracle restatement of the step, torch CPU fp32, all host threads."""
    import torch
    from oracle import dmc_oracle as O
    cfg = CONFIGS[cfg_name]
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sd = O.build_state(cfg['num_class'], cfg['arch_d'], seed=1)
    tr = O.OracleTrainer(sd, O.HParams(), gan=cfg['gan'], arch_d=cfg['arch_d'])
    flow, mv, res, target = O.make_inputs(sample_batch, 3, cfg['num_class'], seed=0)
    per = 2 if cfg['gan'] else 1
    for _ in range(warmup * per):
        tr.step(flow, mv, res, target)
    t0 = time.perf_counter()
    for _ in range(steps * per):
        tr.step(flow, mv, res, target)
    dt = (time.perf_counter() - t0) / (steps * per)
    return sample_batch / dt, dt * 1e3, cores


def run_reference(args):
    rank, _, world = dist_env()
    if rank != 0:
        return
    cfg = CONFIGS[args.config]
    sb = args.cpu_sample_batch
    steps, warmup = min(args.steps, 3), min(args.warmup, 1)
    rate, ms, cores = cpu_reference_rate(args.config, steps, warmup, sb)
    sample = '%d timed + %d warm-up steps of B=%d clips (x3 segments) of the same workload' % (steps, warmup, sb)
    line = {
        'impl': 'reference', 'metric': 'clips/sec', 'value': rate, 'unit': 'clips/s', 'n_gpus': args.gpus,
        'steps': steps, 'warmup': warmup, 'ms_per_step': ms, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': cfg['workload'], 'note': 'CPU, torch fp32, oracle restatement of the reference step'},
        'cpu_baseline': {'value': rate, 'unit': 'clips/s', 'cores': cores, 'kind': 'port', 'sample': sample},
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


def main():
    protect_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--config', default='dmcnet', choices=list(CONFIGS))
    ap.add_argument('--batch', type=int, default=64, help='clips per GPU')
    ap.add_argument('--cpu-sample-batch', type=int, default=4)
    ap.add_argument('--no-graph', action='store_true')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--input', default='fp32', choices=['fp32', 'u8'],
                    help="host format of the e2e leg: three normalised fp32 tensors (default, the measured "
                         "path) or the uint8 [B,S,H,W,7] sample stack normalised on the device")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == 'b200' else args.warmup
    if args.impl == 'reference':
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from dmcnet_b200 import ops
    from dmcnet_b200.engine import DmcEngine
    from dmcnet_b200.trainer import FusedTrainStep, HParams

    rank, local_rank, world = dist_env()
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        import datetime
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank),
                                timeout=datetime.timedelta(seconds=180))
    cfg = CONFIGS[args.config]
    B, S, H, W = args.batch, 3, 224, 224
    per = 2 if cfg['gan'] else 1                      # GAN: one "step" = D-step + G-step pair / 2

    # synthetic inputs with the CoviarDataSet value model (uint8 -> normalised fp32), seeded per rank
    g = torch.Generator().manual_seed(1234 + rank)

    def u8(shape, sigma):
        return torch.clamp(torch.round(128.0 + sigma * torch.randn(shape, generator=g)), 0, 255)
    std = torch.tensor((0.229, 0.224, 0.225))
    mv = ((u8((B, S, 2, H, W), 25.0) / 255.0 - 0.5) / std.mean()).float().pin_memory()
    res = ((u8((B, S, 3, H, W), 20.0) / 255.0 - 0.5) / std.view(1, 1, 3, 1, 1)).float().pin_memory()
    flow = ((u8((B, S, 2, H, W), 30.0) / 255.0 - 0.5) / std.mean()).float().pin_memory()
    target = torch.randint(0, cfg['num_class'], (B,), generator=g).pin_memory()

    # random-init weights of the reference architecture (torchvision resnet18 + reference conv inits)
    from dmcnet_b200.model import build_state
    sd = build_state(cfg['num_class'], cfg['arch_d'], seed=1)
    eng = DmcEngine(cfg['num_class'], S, B * S, gan=cfg['gan'], arch_d=cfg['arch_d'])
    eng.load_state(sd)
    tr = FusedTrainStep(eng, HParams(), B, world_size=world, use_graph=not args.no_graph, pipelined=True)

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
    tr.load_inputs(flow, mv, res, target)
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
    # FusedTrainStep.step_pipelined: every step copies ITS batch from pinned host memory
    # (on a copy stream, overlapping the previous step's compute) and reads back the metrics
    # of the step that just finished; flush() inside the timed region collects the last one.
    stack_u8 = None
    if args.input == 'u8':
        # the same sample values as one interleaved uint8 stack (flow | mv | residual channels)
        denorm = lambda t, d: torch.round((t * d + 0.5) * 255.0).clamp_(0, 255)
        stack_u8 = torch.cat((denorm(flow, std.mean()), denorm(mv, std.mean()),
                              denorm(res, std.view(1, 1, 3, 1, 1))), 2).permute(0, 1, 3, 4, 2)
        stack_u8 = stack_u8.to(torch.uint8).contiguous().pin_memory()

    def e2e_step():
        mk = masks_d if tr._mode() == 'D' else masks_g
        if stack_u8 is not None:
            return tr.step_pipelined_u8(stack_u8, target, masks=mk)
        return tr.step_pipelined(flow, mv, res, target, masks=mk)
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
    sampler.stop_flag = True
    sampler.join(timeout=2)
    h2d = (flow.numel() + mv.numel() + res.numel()) * 4 + target.numel() * 8
    if stack_u8 is not None:
        h2d = stack_u8.numel() + target.numel() * 8
    d2h = 16 * 8                                     # one pinned 16-double stats record per step

    # ---------------- roofline of the dominant kernel family (instrumented eager pass)
    # every rank runs it (the steps contain the gradient all-reduce); rank 0 reports
    from dmcnet_b200.profiling import measure_roofline
    roof = measure_roofline(resident_step, per, min(args.steps, 3), tr)
    barrier()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    clips = B * world
    line = {
        'metric': 'clips/sec', 'value': clips / (ms_step * 1e-3), 'unit': 'clips/s', 'n_gpus': world,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms_step, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32 (bf16x3 split on tensor cores, fp32 accumulate)',
        'data': 'synthetic',
        'config': {'workload': cfg['workload'], 'clips_per_gpu': B, 'global_batch': clips, 'segments': S,
                   'parallelism': 'dp%d' % world,
                   'l2_policy': 'inputs+activations per step (>2 GB) far exceed the 126 MB L2',
                   'cuda_graph': bool(tr.use_graph), 'e2e_input': args.input},
        'e2e': {'value': clips / (ms_e2e * 1e-3), 'unit': 'clips/s', 'h2d_bytes_per_step': h2d,
                'd2h_bytes_per_step': d2h, 'ms_per_step': ms_e2e},
        'gpu_launches': int(launches),
        'clocks': sampler.summary(),
        'algorithmic_tflops': GFLOP_PER_CLIP[args.config] * clips / ms_step,
        'last_metrics': last,
    }
    if roof:
        line['roofline'] = roof.pop('roofline')
        line['kernel_breakdown_ms_per_step'] = roof['breakdown']
    if not args.no_cpu_baseline and world >= 1:
        rate, ms, cores = cpu_reference_rate(args.config, 2, 1, args.cpu_sample_batch)
        line['cpu_baseline'] = {'value': rate, 'unit': 'clips/s', 'cores': cores, 'kind': 'port',
                                'sample': '2 timed + 1 warm-up steps of B=%d clips (x3 segments) of the same '
                                          'workload, torch CPU fp32, oracle restatement' % args.cpu_sample_batch}
    emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
