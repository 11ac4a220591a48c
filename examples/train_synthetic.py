#!/usr/bin/env python
"""End-to-end use of the public API on synthetic data: what ``main()`` of the reference's
train.py does (code/dmcnet/train.py:29-201; GAN: code/dmcnet_GAN/train.py), minus the MPEG-4
loader -- model state, optional ``--weights`` / ``--resume``, the epoch loop with validation and
checkpoints in the reference's file format.

    python examples/train_synthetic.py --epochs 2 --batch-size 8 --model-prefix /tmp/run/hmdb51
    python examples/train_synthetic.py --gan --arch_d Discriminator3 --epochs 1
    torchrun --nproc-per-node 2 --master-addr 127.0.0.1 examples/train_synthetic.py --batch-size 16

Needs a B200 and the built library (``python -c "import __graft_entry__ as g; g.build()"``).
Batches are uint8 ``[B, S, 224, 224, 7]`` stacks (flow | mv | residual), normalised, flipped and
block-averaged on the device exactly as CoviarDataSet / GroupRandomHorizontalFlip would.
"""
import argparse
import os
import random
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np          # noqa: E402
import torch                # noqa: E402


class SyntheticLoader:
    """``batches`` random uint8 sample stacks per epoch, value model of SURVEY.md section 8d."""

    def __init__(self, batches, batch, segments, num_class, seed):
        self.batches, self.B, self.S, self.C, self.seed = batches, batch, segments, num_class, seed

    def __len__(self):
        return self.batches

    def __iter__(self):
        rng = np.random.default_rng(self.seed)
        sig = np.array([30, 30, 25, 25, 20, 20, 20], dtype=np.float32)
        for _ in range(self.batches):
            x = 128 + sig * rng.standard_normal((self.B, self.S, 224, 224, 7), dtype=np.float32)
            stack = torch.from_numpy(np.clip(np.rint(x), 0, 255).astype(np.uint8))
            target = torch.from_numpy(rng.integers(0, self.C, self.B))
            if torch.cuda.is_available():
                stack, target = stack.pin_memory(), target.pin_memory()
            yield stack, target


class U8Step:
    """Adapts the (input_flow, input_mv, input_residual, target) protocol of ``loop.fit`` to uint8
    batches: the loader yields (stack, target); flips are drawn here, one per clip, as
    GroupRandomHorizontalFlip does (code/dmcnet/transforms.py:49)."""

    def __init__(self, step, flow_ds_factor):
        self.step_, self.ds = step, flow_ds_factor

    def __getattr__(self, name):
        return getattr(self.step_, name)

    def step(self, stack, target, *_):
        flips = [random.random() < 0.5 for _ in range(stack.shape[0])]
        return self.step_.step_u8(stack, target, flow_ds_factor=self.ds, flip=flips)

    def validate_batch(self, stack, target, *_):
        self.step_.load_inputs_u8(stack, target, self.ds)
        s = self.step_
        return s.validate_batch(s.in_flow, s.in_mv, s.in_res, s.target)


class _Pairs:
    """(stack, target) -> the 4-tuples ``loop`` unpacks."""

    def __init__(self, loader):
        self.loader = loader

    def __len__(self):
        return len(self.loader)

    def __iter__(self):
        for stack, target in self.loader:
            yield stack, target, None, target


def main():
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument('--data-name', default='hmdb51', choices=['ucf101', 'hmdb51', 'kinetics400'])
    ap.add_argument('--num_segments', type=int, default=3)
    ap.add_argument('--batch-size', type=int, default=8, help='global batch (sharded over the ranks)')
    ap.add_argument('--epochs', type=int, default=2)
    ap.add_argument('--batches-per-epoch', type=int, default=4)
    ap.add_argument('--lr', type=float, default=0.01)
    ap.add_argument('--lr-steps', type=int, nargs='+', default=[20, 35, 45])
    ap.add_argument('--epoch-thre', type=int, default=0)
    ap.add_argument('--eval-freq', type=int, default=1)
    ap.add_argument('--loss-mse', default='MSELoss')
    ap.add_argument('--flow_ds_factor', type=int, default=16)
    ap.add_argument('--arch_estimator', default='DenseNetTiny')
    ap.add_argument('--gan', action='store_true')
    ap.add_argument('--arch_d', default='Discriminator3')
    ap.add_argument('--weights', default=None)
    ap.add_argument('--resume', default=None)
    ap.add_argument('--model-prefix', default=None)
    args = ap.parse_args()

    num_class = {'ucf101': 101, 'hmdb51': 51, 'kinetics400': 400}[args.data_name]      # train.py:43-50
    from dmcnet_b200 import checkpoint, loop
    from dmcnet_b200.engine import DmcEngine
    from dmcnet_b200.model import DENSE_GROWTH, build_state
    from dmcnet_b200.trainer import FusedTrainStep, HParams, shard_range

    rank, world = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1))
    torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', 0)))
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl')
    lo, hi = shard_range(rank, world, args.batch_size)
    B = hi - lo

    arch_d = args.arch_d if args.gan else None
    state = build_state(num_class, arch_d, seed=1, arch_estimator=args.arch_estimator)
    eng = DmcEngine(num_class, args.num_segments, B * args.num_segments, gan=args.gan, arch_d=arch_d,
                    gen_growth=DENSE_GROWTH[args.arch_estimator])
    eng.load_state(state)
    hp = HParams(lr=args.lr, lr_steps=tuple(args.lr_steps), num_segments=args.num_segments, loss_mse=args.loss_mse)
    step = FusedTrainStep(eng, hp, B, world_size=world)
    start_epoch, best = 0, 0.0
    if args.weights:
        missing, unexpected = step.warm_start(checkpoint.load_checkpoint(args.weights)['state_dict'])
        print('warm start: %d missing, %d unexpected keys' % (len(missing), len(unexpected)))
    if args.resume:
        start_epoch, best = step.resume(checkpoint.load_checkpoint(args.resume))
        print("=> loaded checkpoint '{}' (epoch {})".format(args.resume, start_epoch))

    train = _Pairs(SyntheticLoader(args.batches_per_epoch, B, args.num_segments, num_class, seed=10 + rank))
    val = _Pairs(SyntheticLoader(2, B, args.num_segments, num_class, seed=1000 + rank))
    log = print if rank == 0 else (lambda *_: None)
    best = loop.fit(U8Step(step, args.flow_ds_factor), train, val, epochs=args.epochs, start_epoch=start_epoch,
                    best_prec1=best, eval_freq=args.eval_freq, epoch_thre=args.epoch_thre, gan=args.gan,
                    segments=args.num_segments, model_prefix=args.model_prefix if rank == 0 else None, log=log)
    log('best Prec@1 %.3f' % best)
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
