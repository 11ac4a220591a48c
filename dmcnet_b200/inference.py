"""Video-level scoring (the deployment side of the same kernels).

Mirrors the protocol of code/dmcnet/test.py:88-198 (GAN flavour
code/dmcnet_GAN/test.py): one video = ``test_segments x test_crops`` frames
(25 x 10 = 250 in the shipped recipe), eval-mode forward (BatchNorm running
statistics, no dropout; the discriminator only on request), class scores = mean of the LOGITS over
all frames of the video (test.py:147-148), accuracy = argmax against the label
(test.py:173-179), and the ``--save-scores`` ``.npz`` consumed by
code/dmcnet/combine.py:35-56 (``scores`` = one ``(scores[1,C], label)`` pair per
video in sorted-name order, ``labels``, ``names``).

The forward runs on ``DmcEngine`` (hand-written kernels only); the mean over the
frames of a video, the video-level cross-entropy and the top-1/top-5 hits come from
the same ``ce_head`` launch the train step uses (S = frames per video).  There is
no CPU path: ``VideoScorer`` needs a CUDA device and the built library.
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import numpy as np
import torch


# ------------------------------------------------------------------ host logic (no device work)
def strip_module_prefix(state: Dict[str, torch.Tensor]) -> "OrderedDict[str, torch.Tensor]":
    """Checkpoints are saved from the nn.DataParallel wrapper; test.py:84 drops the first
    dotted component of every key (``module.base_model...`` -> ``base_model...``)."""
    return OrderedDict((k.split('.', 1)[1] if '.' in k else '', v) for k, v in state.items())


def plan_launches(frames_per_video: int, max_frames: Optional[int]) -> Tuple[int, int]:
    """(frames per engine launch, launches per video).  The engine runs a fixed frame count,
    so a video is cut into equal parts: the largest divisor of ``frames_per_video`` that
    does not exceed ``max_frames`` (None = the whole video in one launch)."""
    if frames_per_video <= 0:
        raise ValueError('a video needs at least one frame')
    if max_frames is None or max_frames >= frames_per_video:
        return frames_per_video, 1
    if max_frames <= 0:
        raise ValueError('max_frames must be positive')
    per = max(d for d in range(1, max_frames + 1) if frames_per_video % d == 0)
    return per, frames_per_video // per


def check_crops(test_crops: int) -> int:
    """test.py:88-98: only the centre crop or GroupOverSample's 10 crops exist."""
    if test_crops not in (1, 10):
        raise ValueError('Only 1 and 10 crops are supported, but got {}.'.format(test_crops))
    return test_crops


def video_accuracy(output: Sequence[tuple]) -> float:
    """Per-cent of videos whose argmax score equals the label (test.py:173-179)."""
    if not output:
        raise ValueError('no videos were scored')
    hits = sum(int(np.argmax(o[0])) == int(o[1]) for o in output)
    return 100.0 * hits / len(output)


def adversarial_accuracy(output: Sequence[tuple]) -> float:
    """'Accuracy adv G' of code/dmcnet_GAN/test.py:125,131-133: np.argmax over the FLATTENED
    [frames, 2] validity logits of each video, summed and divided by the video count (the
    reference prints this figure as is; it is not bounded by 100)."""
    if not output:
        raise ValueError('no videos were scored')
    return 100.0 * float(sum(int(np.argmax(o[2])) for o in output)) / len(output)


def ordered_for_save(output: Sequence[tuple], names: Sequence[str]):
    """Sorted-name order of the saved file (test.py:183-195).  Returns (scores, labels, names)
    with ``scores`` an object array [n, 2]: column 0 the [1, C] float32 scores, column 1 the
    label -- the element layout combine.py indexes as ``score[0][0]`` / ``score[1]``.  The GAN
    script appends the per-frame discriminator logits as a third element
    (code/dmcnet_GAN/test.py:116): 3-tuples give an [n, 3] array with that column."""
    if len(output) != len(names):
        raise ValueError('%d scored videos but %d names' % (len(output), len(names)))
    if len(set(names)) != len(names):
        raise ValueError('video names must be unique (they key the saved order)')
    width = {len(o) for o in output}
    if len(width) > 1 or not width <= {2, 3}:
        raise ValueError('every entry must be (scores, label) or (scores, label, validity)')
    cols = width.pop() if width else 2
    order = sorted(range(len(names)), key=lambda i: names[i])
    scores = np.empty((len(order), cols), dtype=object)
    for row, i in enumerate(order):
        scores[row, 0] = np.asarray(output[i][0], dtype=np.float32).reshape(1, -1)
        scores[row, 1] = int(output[i][1])
        if cols == 3:
            scores[row, 2] = np.asarray(output[i][2], dtype=np.float32)
    labels = np.array([int(output[i][1]) for i in order], dtype=np.int64)
    return scores, labels, np.array([names[i] for i in order])


def save_scores(path: str, output: Sequence[tuple], names: Sequence[str]) -> None:
    """``--save-scores`` (test.py:181-198)."""
    scores, labels, ordered = ordered_for_save(output, names)
    np.savez(path, scores=scores, labels=labels, names=ordered)


def load_scores(path: str):
    """(scores [n, C] float32, labels [n], names [n]) of a file written by ``save_scores`` or by the
    reference's test.py (object arrays: ``allow_pickle``)."""
    with np.load(path, allow_pickle=True) as z:
        scores = np.stack([np.asarray(s[0], dtype=np.float32)[0] for s in z['scores']])
        labels = np.array([int(s[1]) for s in z['scores']], dtype=np.int64)
        return scores, labels, np.array(z['names'])


def combine_scores(files: Sequence[str], weights: Sequence[float]) -> Tuple[float, int]:
    """Late fusion of per-stream score files (code/dmcnet/combine.py:35-56): weighted sum of
    the video scores, argmax accuracy as a fraction, and the video count."""
    if len(files) != len(weights) or not files:
        raise ValueError('one weight per score file')
    total, ref_labels = None, None
    for f, w in zip(files, weights):
        s, l, _ = load_scores(f)
        if ref_labels is None:
            total, ref_labels = w * s, l
        else:
            assert np.array_equal(l, ref_labels), 'score files disagree on the labels'
            total = total + w * s
    return float(np.mean(np.argmax(total, axis=1) == ref_labels)), len(ref_labels)


# ------------------------------------------------------------------ device path
class VideoScorer:
    """``forward_video`` of test.py:139-151 on the B200 engine.

    state:   state_dict of a trained ``Model`` (``module.`` prefix already stripped; keys of a
             discriminator or ``data_bn`` are ignored, as ``load_state_dict(strict=False)`` does)
    """

    def __init__(self, state: Dict[str, torch.Tensor], num_class: int, test_segments: int = 25,
                 test_crops: int = 10, *, gen_flow_or_delta: int = 1, height: int = 224,
                 width: int = 224, max_frames_per_launch: Optional[int] = None,
                 arch_d: Optional[str] = None, device: Optional[torch.device] = None,
                 arch_estimator: str = 'DenseNetTiny', att: int = 0, gen_flow_ds_factor: int = 0):
        from .engine import DmcEngine
        self.num_class = num_class
        self.arch_d = arch_d
        self.segments, self.crops = test_segments, check_crops(test_crops)
        self.frames = test_segments * test_crops
        self.per_launch, self.launches = plan_launches(self.frames, max_frames_per_launch)
        # arch_d: also run the discriminator on the generated map (code/dmcnet_GAN/test.py:91,
        # validity [frames, 2] per video); the class scores do not depend on it
        # --arch_estimator / --att / --gen_flow_ds_factor of test.py (test_options mirrors train_options)
        from .model import DENSE_GROWTH
        self.eng = DmcEngine(num_class, test_segments, self.per_launch, gan=arch_d is not None,
                             arch_d=arch_d, gen_flow_or_delta=gen_flow_or_delta, height=height,
                             width=width, device=device, arch_estimator=arch_estimator, att=att,
                             gen_flow_ds_factor=gen_flow_ds_factor,
                             gen_growth=DENSE_GROWTH.get(arch_estimator, DENSE_GROWTH['DenseNetTiny']))
        # num_batches_tracked is absent from torch < 0.4.1 checkpoints (the reference's, README.md:28)
        missing = [k for k in list(self.eng.specs) + list(self.eng.buffers)
                   if k not in state and not k.endswith('.num_batches_tracked')]
        if missing:
            raise KeyError('state_dict lacks %d tensors of the scoring path, e.g. %s'
                           % (len(missing), missing[0]))
        self.eng.load_state(state)
        dev = self.eng.device
        self.H, self.W = height, width
        self._logits = torch.zeros(self.frames, num_class, dtype=torch.float32, device=dev)
        self._validity = torch.zeros(self.frames, 2, dtype=torch.float32, device=dev)
        self._scores = torch.zeros(1, num_class, dtype=torch.float32, device=dev)
        self._stats = torch.zeros(4, dtype=torch.float32, device=dev)
        self._label = torch.zeros(1, dtype=torch.int64, device=dev)

    def forward_video(self, input_mv: torch.Tensor, input_residual: torch.Tensor,
                      label: int = 0) -> np.ndarray:
        """[1, frames, 2, H, W] / [1, frames, 3, H, W] (one DataLoader item, batch_size=1 as in
        test.py:119) -> scores [1, num_class] (numpy, like ``scores.data.cpu().numpy()``)."""
        from . import ops
        H, W = self.H, self.W
        mv = input_mv.reshape(-1, 2, H, W)
        res = input_residual.reshape(-1, 3, H, W)
        if mv.shape[0] != self.frames or res.shape[0] != self.frames:
            raise ValueError('expected %d frames per video (test_segments x test_crops), got %d / %d'
                             % (self.frames, mv.shape[0], res.shape[0]))
        dev = self.eng.device
        mv = mv.to(dev, torch.float32, non_blocking=True).contiguous()
        res = res.to(dev, torch.float32, non_blocking=True).contiguous()
        n = self.per_launch
        for j in range(self.launches):
            out = self.eng.forward(mv[j * n:(j + 1) * n], res[j * n:(j + 1) * n], train=False)
            self._logits[j * n:(j + 1) * n].copy_(out[0])
            if self.arch_d is not None:
                self._validity[j * n:(j + 1) * n].copy_(out[1])
        self._label.fill_(int(label))
        # mean over every frame of the video + video-level CE / top-k in one launch
        ops.ce_head(self._logits, 1, self.frames, self.num_class, self._label, 0.0, self._scores, None,
                    self._stats)
        return self._scores.cpu().numpy().copy()

    def forward_video_u8(self, decoded_u8: torch.Tensor, label: int = 0, scale_size: int = 256) -> np.ndarray:
        """The test transform on the device as well: ``decoded_u8`` is the uint8 stack
        [test_segments, Hs, Ws, 7] of the decoded frames (flow | mv | residual channels,
        code/dmcnet/dataset.py:210); with 10 crops it goes through GroupOverSample(crop, scale)
        (five windows of the frame scaled to scale_size x scale_size, each also flipped,
        code/dmcnet/transforms.py:78-114), with 1 crop through GroupScale + GroupCenterCrop
        (test.py:88-98); then the sample arithmetic of dataset.py:215-263 and ``forward_video``.
        The frame ORDER differs from the reference's list (window, segment, flip), which the mean
        over all frames does not see."""
        from . import input_stage as S
        segs, H, W = self.segments, self.H, self.W
        if decoded_u8.dim() != 4 or decoded_u8.shape[0] != segs or decoded_u8.shape[-1] != 7:
            raise ValueError('expected a uint8 stack [%d, Hs, Ws, 7], got %s' % (segs, tuple(decoded_u8.shape)))
        Hs, Ws = int(decoded_u8.shape[1]), int(decoded_u8.shape[2])
        key = (Hs, Ws, scale_size)
        if getattr(self, '_u8_key', None) != key:
            dev = self.eng.device
            self._crop = S.CropResizeStage(segs, Hs, Ws, H, W, device=dev)
            self._norm = S.U8InputStage(segs, H, W, device=dev)
            if self.crops == 10:
                wins = S.oversample_offsets(scale_size, scale_size, H, W)
            else:
                wins = [((scale_size - H) // 2, (scale_size - W) // 2)]
            self._tabs = [S.scaled_crop_tables(Hs, Ws, scale_size, scale_size, r0, c0, H, W)[None] for r0, c0 in wins]
            f32 = dict(dtype=torch.float32, device=dev)
            self._mv_all = torch.empty(self.frames, 2, H, W, **f32)
            self._res_all = torch.empty(self.frames, 3, H, W, **f32)
            self._flow_tmp = torch.empty(segs, 2, H, W, **f32)
            self._u8_key = key
        k = 0
        decoded_u8 = decoded_u8.to(self.eng.device, non_blocking=True)       # one host->device copy
        for tab in self._tabs:
            stack = self._crop(decoded_u8, tab)
            for flipped in ((False, True) if self.crops == 10 else (False,)):
                sl = slice(k * segs, (k + 1) * segs)
                self._norm(stack, self._flow_tmp, self._mv_all[sl], self._res_all[sl],
                           flip=[True] if flipped else None)
                k += 1
        assert k * segs == self.frames
        return self.forward_video(self._mv_all, self._res_all, label)

    def last_validity(self) -> np.ndarray:
        """Discriminator logits [frames, 2] of the last scored video (needs ``arch_d``); the third
        element of the GAN script's output tuples (code/dmcnet_GAN/test.py:97,116)."""
        if self.arch_d is None:
            raise RuntimeError('VideoScorer was built without arch_d: no discriminator was run')
        return self._validity.cpu().numpy().copy()

    def last_stats(self) -> Dict[str, float]:
        """Cross-entropy and top-1 / top-5 hit of the last scored video against its label."""
        s = self._stats.cpu().tolist()
        return {'loss': s[0], 'top1': s[1], 'top5': s[2]}

    def evaluate(self, samples: Iterable) -> List[tuple]:
        """The loop of test.py:155-171 over ``(input_flow, input_mv, input_residual, label)`` items."""
        output = []
        for _flow, mv, res, label in samples:
            lab = int(label[0]) if hasattr(label, '__len__') else int(label)
            scores = self.forward_video(mv, res, lab)
            output.append((scores, lab) if self.arch_d is None else (scores, lab, self.last_validity()))
        return output
