"""TEST INFRASTRUCTURE (checker only; never imported by the product path).

CPU restatement of the video-level testing protocol around ``Model.forward``:
  * forward_video        code/dmcnet/test.py:139-151
  * accuracy             code/dmcnet/test.py:173-179
  * --save-scores        code/dmcnet/test.py:181-198
  * late fusion          code/dmcnet/combine.py:35-56
The model forward itself is ``oracle.dmc_oracle.model_forward`` (pinned against the
reference's model.py).  test.py cannot be imported here (``async=True`` is a
SyntaxError on Python >= 3.7), so these few lines are restated.
Pins: the score-file layout and the fusion formula are checked against files the reference's
own test.py wrote (exp_my/*/split*/*_score_model_best.npz; fixture tests/golden/score_files.npz,
and the published 64.05 / 61.31 / 60.07 % on the full files); ``forward_video`` and ``accuracy``
are a view + mean / argmax over the pinned forward -- for those two lines themselves: parity
unpinned (no reference output exists without trained weights and videos).
"""
from typing import Dict, List, Sequence, Tuple

import numpy as np
import torch

from . import dmc_oracle as O


def forward_video(state: Dict[str, torch.Tensor], input_mv: torch.Tensor, input_residual: torch.Tensor,
                  test_segments: int, test_crops: int) -> np.ndarray:
    """test.py:139-151: eval forward, view(-1, segments*crops, C), mean over dim 1."""
    st = {k: v.detach() for k, v in state.items()}
    with torch.no_grad():
        scores, _ = O.model_forward(st, input_mv, input_residual, train=False)
    scores = scores.view((-1, test_segments * test_crops) + tuple(scores.shape[1:]))
    return torch.mean(scores, dim=1).numpy().copy()


def forward_video_gan(state: Dict[str, torch.Tensor], input_mv: torch.Tensor, input_residual: torch.Tensor,
                      test_segments: int, test_crops: int, arch_d: str):
    """code/dmcnet_GAN/test.py:86-98: (scores [1,C], validity [frames,2])."""
    st = {k: v.detach() for k, v in state.items()}
    with torch.no_grad():
        scores, validity, _ = O.model_forward(st, input_mv, input_residual, None, gan=True, arch_d=arch_d,
                                              train=False)
    scores = scores.view((-1, test_segments * test_crops) + tuple(scores.shape[1:]))
    return torch.mean(scores, dim=1).numpy().copy(), validity.numpy().copy()


def accuracy(output: Sequence[Tuple[np.ndarray, int]]) -> float:
    """test.py:173-179."""
    video_pred = [np.argmax(x[0]) for x in output]
    video_labels = [x[1] for x in output]
    return float(np.sum(np.array(video_pred) == np.array(video_labels))) / len(video_pred) * 100.0


def save_scores(path: str, output: Sequence[Tuple[np.ndarray, int]], name_list: List[str]) -> None:
    """test.py:181-198.  numpy < 1.24 turned the list of (ndarray[1,C], label) tuples into an
    object array implicitly; the explicit dtype=object below is that behaviour."""
    video_labels = [x[1] for x in output]
    order_dict = {e: i for i, e in enumerate(sorted(name_list))}
    reorder_output = [None] * len(output)
    reorder_label = [None] * len(output)
    reorder_name = [None] * len(output)
    for i in range(len(output)):
        idx = order_dict[name_list[i]]
        reorder_output[idx] = output[i]
        reorder_label[idx] = video_labels[i]
        reorder_name[idx] = name_list[i]
    scores = np.empty((len(output), 2), dtype=object)
    for i, (s, l) in enumerate(reorder_output):
        scores[i, 0], scores[i, 1] = s, l
    np.savez(path, scores=scores, labels=reorder_label, names=reorder_name)


def combine(files: Sequence[str], weights: Sequence[float]) -> Tuple[float, int]:
    """combine.py:35-56 for any number of streams (iframe/mv/res[/flow] in the reference)."""
    loaded = [np.load(f, allow_pickle=True) for f in files]
    n = len(loaded[0]['names'])
    per_stream = [np.array([score[0][0] for score in z['scores']]) for z in loaded]
    labels = [np.array([score[1] for score in z['scores']]) for z in loaded]
    for l in labels[1:]:
        assert np.all(labels[0] == l)
    combined_score = sum(w * s for w, s in zip(weights, per_stream))
    accuracy_ = float(sum(np.argmax(combined_score, axis=1) == labels[0])) / n
    return accuracy_, n
