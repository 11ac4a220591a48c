"""Checkpoint / resume in the reference's own file format.

The reference saves ``{'epoch', 'arch', 'state_dict', 'best_prec1', 'optimizer_cls',
'optimizer_gf'[, 'optimizer_d']}`` with ``torch.save`` (code/dmcnet/train.py:190-201,
:372-377; GAN code/dmcnet_GAN/train.py:203-215, :488-493): ``state_dict`` comes from the
``nn.DataParallel`` wrapper (every key prefixed ``module.``) and each ``optimizer_*`` is
the ``state_dict()`` of a ``torch.optim.Adam`` built with ONE param group per tensor
(train.py:121-142).  ``--resume`` restores all of it (train.py:145-163); ``--weights``
warm-starts the model only, ``strict=False`` after dropping the prefix (train.py:64-68)
-- that is how the GAN stage starts from the flow-reconstruction stage.

Here the optimizer state lives in flat fp32 buckets (first/second moments) with one
device step counter per optimizer.  The functions below translate between the two
representations, so a checkpoint written by either side resumes on the other.  They work
on any "bucket owner" exposing ``specs`` (name -> shape, in ``named_parameters`` order),
``offsets``, ``exp_avg``, ``exp_avg_sq`` -- the engine on the GPU, a plain stand-in in the
CPU tests.
"""
from __future__ import annotations

import os
import shutil
from collections import OrderedDict
from typing import Dict, List, Optional, Sequence, Tuple

import torch

GROUP_TAGS = ('base_model', 'gen_flow_model', 'discriminator')
OPTIMIZER_KEYS = {'base_model': 'optimizer_cls', 'gen_flow_model': 'optimizer_gf',
                  'discriminator': 'optimizer_d'}


def group_keys(specs, tag: str) -> List[str]:
    """Parameters of one optimizer, in the order the reference appends their groups
    (substring test on the name, code/dmcnet/train.py:125,129; GAN :135)."""
    return [k for k in specs if tag in k]


def _numel(shape) -> int:
    n = 1
    for d in shape:
        n *= int(d)
    return n


# ------------------------------------------------------------------ optimizer state
def adam_state_to_torch(owner, tag: str, step: int, hyper_rows: Dict[str, Tuple[float, float]],
                        lr_mult: float, betas=(0.9, 0.999), eps: float = 1e-3) -> dict:
    """``torch.optim.Adam.state_dict()`` of the optimizer that owns the ``tag`` parameters.
    hyper_rows: name -> (current lr, current weight_decay) of its group (what
    ``adjust_learning_rate`` last wrote, train.py:398-408).  No state entries exist before the
    first step, as in torch."""
    keys = group_keys(owner.specs, tag)
    state, groups = {}, []
    for i, k in enumerate(keys):
        lr, wd = hyper_rows[k]
        groups.append({'lr': float(lr), 'betas': tuple(betas), 'eps': float(eps), 'weight_decay': float(wd),
                       'amsgrad': False, 'lr_mult': float(lr_mult),
                       'decay_mult': 0.0 if 'bias' in k else 1.0, 'params': [i]})
        if step > 0:
            o, n, shp = owner.offsets[k], _numel(owner.specs[k]), tuple(owner.specs[k])
            state[i] = {'step': torch.tensor(float(step)),
                        'exp_avg': owner.exp_avg[o:o + n].detach().reshape(shp).cpu().clone(),
                        'exp_avg_sq': owner.exp_avg_sq[o:o + n].detach().reshape(shp).cpu().clone()}
    return {'state': state, 'param_groups': groups}


def adam_state_from_torch(owner, tag: str, opt_state: dict) -> int:
    """Load a ``torch.optim.Adam.state_dict()`` (one group per tensor, reference order) into the
    flat moment buckets; returns the step count (0 when the optimizer never stepped).  Raises if
    the checkpoint's groups do not match this model's parameters."""
    keys = group_keys(owner.specs, tag)
    groups = opt_state['param_groups']
    flat_ids = [pid for g in groups for pid in g['params']]
    if len(flat_ids) != len(keys):
        raise ValueError('%s: checkpoint has %d parameters, the model has %d'
                         % (OPTIMIZER_KEYS[tag], len(flat_ids), len(keys)))
    state = opt_state['state']
    steps = set()
    for pid, k in zip(flat_ids, keys):
        o, n = owner.offsets[k], _numel(owner.specs[k])
        st = state.get(pid, state.get(str(pid)))
        if st is None:                                   # never stepped (or frozen): zero moments
            owner.exp_avg[o:o + n].zero_()
            owner.exp_avg_sq[o:o + n].zero_()
            steps.add(0)
            continue
        if tuple(st['exp_avg'].shape) != tuple(owner.specs[k]):
            raise ValueError('%s: %s has shape %s in the checkpoint, %s in the model'
                             % (OPTIMIZER_KEYS[tag], k, tuple(st['exp_avg'].shape), tuple(owner.specs[k])))
        owner.exp_avg[o:o + n].copy_(st['exp_avg'].reshape(-1).to(owner.exp_avg.dtype))
        owner.exp_avg_sq[o:o + n].copy_(st['exp_avg_sq'].reshape(-1).to(owner.exp_avg_sq.dtype))
        steps.add(int(st['step']))                       # int (torch 0.4) or 0-dim tensor (torch >= 1.12)
    if len(steps) > 1:
        raise ValueError('%s: parameters disagree on the step count %s (one counter per optimizer here)'
                         % (OPTIMIZER_KEYS[tag], sorted(steps)))
    return steps.pop() if steps else 0


# ------------------------------------------------------------------ model state
def add_module_prefix(state: Dict[str, torch.Tensor]) -> "OrderedDict[str, torch.Tensor]":
    """Keys as ``nn.DataParallel(model).state_dict()`` writes them (train.py:117, :193)."""
    return OrderedDict(('module.' + k, v) for k, v in state.items())


def strip_first_component(state: Dict[str, torch.Tensor]) -> "OrderedDict[str, torch.Tensor]":
    """``'.'.join(k.split('.')[1:])`` of train.py:67 / test.py:84."""
    return OrderedDict(('.'.join(k.split('.')[1:]), v) for k, v in state.items())


def merge_non_strict(current: Dict[str, torch.Tensor], loaded: Dict[str, torch.Tensor]
                     ) -> Tuple["OrderedDict[str, torch.Tensor]", List[str], List[str]]:
    """``load_state_dict(loaded, strict=False)`` on a model whose state is ``current``: tensors
    present in both are taken from ``loaded`` (shape mismatch is an error, as in torch); returns
    (merged, missing_keys, unexpected_keys)."""
    merged = OrderedDict()
    for k, v in current.items():
        if k in loaded:
            if tuple(loaded[k].shape) != tuple(v.shape):
                raise RuntimeError('size mismatch for %s: copying a param with shape %s from checkpoint, '
                                   'the shape in current model is %s.'
                                   % (k, tuple(loaded[k].shape), tuple(v.shape)))
            merged[k] = loaded[k]
        else:
            merged[k] = v
    # torch's own loader does not report a missing num_batches_tracked (BN version < 2 checkpoints)
    missing = [k for k in current if k not in loaded and not k.endswith('.num_batches_tracked')]
    unexpected = [k for k in loaded if k not in current]
    return merged, missing, unexpected


# ------------------------------------------------------------------ files
def checkpoint_names(model_prefix: str, representation: str, filename: str = 'checkpoint.pth.tar'
                     ) -> Tuple[str, str]:
    """(<prefix>_<repr>_checkpoint.pth.tar, <prefix>_<repr>_model_best.pth.tar), train.py:372-377."""
    rep = representation.lower()
    return '_'.join((model_prefix, rep, filename)), '_'.join((model_prefix, rep, 'model_best.pth.tar'))


def save_checkpoint(state: dict, is_best: bool, model_prefix: str, representation: str,
                    filename: str = 'checkpoint.pth.tar') -> str:
    """``save_checkpoint`` of train.py:372-377."""
    path, best = checkpoint_names(model_prefix, representation, filename)
    d = os.path.dirname(path)
    if d:
        os.makedirs(d, exist_ok=True)
    torch.save(state, path)
    if is_best:
        shutil.copyfile(path, best)
    return path


def load_checkpoint(path: str) -> dict:
    """``torch.load(path, map_location=cpu)`` (train.py:66, :148); the files hold plain tensors,
    dicts, tuples and floats."""
    return torch.load(path, map_location='cpu', weights_only=False)
