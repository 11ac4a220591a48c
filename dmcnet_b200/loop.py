"""Epoch driver around the fused step: what ``main`` / ``train`` / ``validate`` of the
reference scripts do between batches (code/dmcnet/train.py:173-201, :205-293, :296-370;
code/dmcnet_GAN/train.py:186-215, :218-400, :403-486).

It owns nothing of the hot path: every batch goes to ``FusedTrainStep.step`` /
``validate_batch``; this module keeps the running averages, the print cadence, the
epoch schedule (``set_epoch`` = ``adjust_learning_rate``), the "validate every
``eval_freq`` epochs and on the last one" rule and the "save when best or every
``SAVE_FREQ`` epochs" rule, and writes checkpoints in the reference's format
(``checkpoint.py``).  Loaders are any iterable of ``(input_flow, input_mv, input_residual,
target)`` batches in the CoviarDataSet layout.
"""
from __future__ import annotations

import time
from typing import Callable, Dict, Iterable, Optional

SAVE_FREQ = 40           # code/dmcnet/train.py:26
PRINT_FREQ = {False: 20, True: 15}        # train.py:27 / GAN train.py:27


class AverageMeter:
    """Running value / average (code/dmcnet/train.py:380-395)."""

    def __init__(self):
        self.reset()

    def reset(self):
        self.val = 0
        self.avg = 0
        self.sum = 0
        self.count = 0

    def update(self, val, n=1):
        self.val = val
        self.sum += val * n
        self.count += n
        self.avg = self.sum / self.count


def _fmt(meters: Dict[str, AverageMeter], names) -> str:
    return '\t'.join('{} {:.4f} ({:.4f})'.format(k, meters[k].val, meters[k].avg) for k in names if k in meters)


def train_epoch(step, loader: Iterable, epoch: int, *, gan: bool = False, segments: int = 3,
                log: Callable[[str], None] = print) -> Dict[str, float]:
    """``train`` for one epoch.  Meters are weighted by frames (``input_flow.size(0)`` after the
    view of train.py:230, i.e. B*S).  GAN: a batch is a D-step or a G-step by parity of the
    iteration; metrics of the two kinds are averaged separately (GAN/train.py:303-311, :361-366)."""
    meters: Dict[str, AverageMeter] = {}
    batch_time, data_time = AverageMeter(), AverageMeter()
    end = time.time()
    n_batches = len(loader) if hasattr(loader, '__len__') else -1
    for i, (input_flow, input_mv, input_residual, target) in enumerate(loader):
        data_time.update(time.time() - end)
        m = step.step(input_flow, input_mv, input_residual, target)
        frames = int(target.shape[0]) * segments
        kind = ('D_' if 'loss_mse' not in m else 'G_') if gan else ''
        for k, v in m.items():
            meters.setdefault(kind + k, AverageMeter()).update(v, frames)
        batch_time.update(time.time() - end)
        end = time.time()
        if i % PRINT_FREQ[gan] == 0:
            log('Epoch: [{0}][{1}/{2}]\tTime {3:.3f} ({4:.3f})\tData {5:.3f} ({6:.3f})\t{7}'.format(
                epoch, i, n_batches, batch_time.val, batch_time.avg, data_time.val, data_time.avg,
                _fmt(meters, sorted(meters))))
    return {k: v.avg for k, v in meters.items()}


def validate_epoch(step, loader: Iterable, *, gan: bool = False, segments: int = 3,
                   log: Callable[[str], None] = print) -> Dict[str, float]:
    """``validate``: eval-mode batches, frame-weighted averages; returns them (``prec1`` is what
    ``main`` compares with ``best_prec1``)."""
    meters: Dict[str, AverageMeter] = {}
    n_batches = len(loader) if hasattr(loader, '__len__') else -1
    for i, (input_flow, input_mv, input_residual, target) in enumerate(loader):
        m = step.validate_batch(input_flow, input_mv, input_residual, target)
        frames = int(target.shape[0]) * segments
        for k, v in m.items():
            meters.setdefault(k, AverageMeter()).update(v, frames)
        if i % PRINT_FREQ[gan] == 0:
            log('Test: [{0}/{1}]\t{2}'.format(i, n_batches, _fmt(meters, sorted(meters))))
    out = {k: v.avg for k, v in meters.items()}
    log('Testing Results: Prec@1 {:.3f} Prec@5 {:.3f} Loss {:.5f}'.format(
        out.get('prec1', 0.0), out.get('prec5', 0.0), out.get('loss', 0.0)))
    return out


def fit(step, train_loader: Iterable, val_loader: Optional[Iterable], *, epochs: int, start_epoch: int = 0,
        best_prec1: float = 0.0, eval_freq: int = 5, epoch_thre: int = 0, gan: bool = False,
        segments: int = 3, arch: str = 'resnet18', model_prefix: Optional[str] = None,
        representation: str = 'mv', log: Callable[[str], None] = print) -> float:
    """The epoch loop of ``main`` (train.py:173-201).  Returns the best validation Prec@1."""
    from . import checkpoint as C
    for epoch in range(start_epoch, epochs):
        step.set_epoch(epoch, epoch_thre=epoch_thre)                          # :175-176
        if not gan:
            log('current epoch freeze?: {}'.format(str(epoch < epoch_thre)))  # :179
        train_epoch(step, train_loader, epoch, gan=gan, segments=segments, log=log)
        if val_loader is not None and (epoch % eval_freq == 0 or epoch == epochs - 1):   # :186
            prec1 = validate_epoch(step, val_loader, gan=gan, segments=segments, log=log)['prec1']
            is_best = prec1 > best_prec1
            best_prec1 = max(prec1, best_prec1)
            if model_prefix is not None and (is_best or epoch % SAVE_FREQ == 0):          # :190
                C.save_checkpoint(step.checkpoint(epoch + 1, arch, best_prec1), is_best, model_prefix,
                                  representation)
    return best_prec1
