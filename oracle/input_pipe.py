"""TEST INFRASTRUCTURE (checker only; never imported by the product path).

CPU restatement of the tail of ``CoviarDataSet.__getitem__``
(code/dmcnet/dataset.py:215-263, identical in code/dmcnet_GAN/dataset.py): from the
list of augmented uint8 ``[H, W, 7]`` stacks (flow x,y | mv x,y | residual r,g,b;
dataset.py:210) to the three normalised float tensors the model consumes.

Third-party arithmetic not under /root/reference: ``skimage.measure.block_reduce``
(scikit-image, version unpinned by the reference, not installed here).  Its published
algorithm is restated in ``block_reduce``: pad every axis at its END with ``cval`` (0) up
to a multiple of the block size, view as blocks, apply ``func`` over the block axes.

Pinned (``oracle/pin_input_pipe.py``, ``tests/test_oracle_pin.py``): bit-identical to the
reference's own ``dataset.py`` executed here with stub ``coviar`` / ``skimage`` modules on
synthetic decoded frames; the reference's outputs are committed as
``tests/golden/input_pipe.npz``.  ``block_reduce`` itself is "parity unpinned" (restated
from the documentation) -- the factors used in practice (16 | 224) never pad.
"""
from typing import Sequence, Tuple

import numpy as np
import torch

INPUT_STD = (0.229, 0.224, 0.225)       # code/dmcnet/dataset.py:111-112


def block_reduce(image: np.ndarray, block_size: Sequence[int], func=np.mean, cval=0) -> np.ndarray:
    """skimage.measure.block_reduce (restated, see the module docstring)."""
    assert len(block_size) == image.ndim
    pad = [(0, (-image.shape[i]) % block_size[i]) for i in range(image.ndim)]
    image = np.pad(image, pad, mode='constant', constant_values=cval)
    shape = []
    for i in range(image.ndim):
        shape += [image.shape[i] // block_size[i], block_size[i]]
    blocked = image.reshape(shape)
    return func(blocked, axis=tuple(range(1, 2 * image.ndim, 2)))


def flip_group(img_group: Sequence[np.ndarray]):
    """The flipped branch of GroupRandomHorizontalFlip.__call__ (code/dmcnet/transforms.py:49-57):
    mirror, then negate the x components of flow (channel 0) and mv (channel 2) around 128 in int32.
    Note v = 0 -> 256: the result no longer fits uint8."""
    ret = [img[:, ::-1, :].astype(np.int32) for img in img_group]
    for i in range(len(ret)):
        ret[i][:, :, :4] -= 128
        ret[i][..., 0] *= (-1)
        ret[i][..., 2] *= (-1)
        ret[i][:, :, :4] += 128
    return ret


def sample_from_frames(frames: Sequence[np.ndarray], flow_ds_factor: int = 0
                       ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """dataset.py:215-263 for representation == 'mv', upsample_interp == False.
    frames: list of uint8 [H, W, 7] -> (input_flow [S,2,H,W], input_mv [S,2,H,W],
    input_residual [S,3,H,W]) float32."""
    frames = np.array(frames)                                  # :217
    frames = np.transpose(frames, (0, 3, 1, 2))                # :218
    input_flow = frames[:, 0:2, :, :]                          # :222-224
    input_mv = frames[:, 2:4, :, :]
    input_residual = frames[:, 4:, :, :]
    if flow_ds_factor != 0:                                    # :226-246
        factor = flow_ds_factor
        w_max = input_flow.shape[2]
        h_max = input_flow.shape[3]
        input_flow = block_reduce(input_flow, block_size=(1, 1, factor, factor), func=np.mean)
        input_flow = input_flow.repeat(factor, axis=2).repeat(factor, axis=3)
        input_flow = input_flow[:, :, :w_max, :h_max]
    input_std = torch.from_numpy(np.array(INPUT_STD).reshape((1, 3, 1, 1))).float()   # :111-112
    input_flow = torch.from_numpy(np.ascontiguousarray(input_flow)).float() / 255.0   # :251-253
    input_mv = torch.from_numpy(np.ascontiguousarray(input_mv)).float() / 255.0
    input_residual = torch.from_numpy(np.ascontiguousarray(input_residual)).float() / 255.0
    input_mv = (input_mv - 0.5) / torch.mean(input_std)        # :260
    input_flow = (input_flow - 0.5) / torch.mean(input_std)    # :262
    input_residual = (input_residual - 0.5) / input_std        # :263
    return input_flow, input_mv, input_residual


def synthetic_frames(segments: int, height: int, width: int, seed: int = 0) -> np.ndarray:
    """uint8 [S, H, W, 7] with the value model of SURVEY.md section 8d (flow 128 +- 30, mv 128 +- 25,
    residual 128 +- 20, clipped), plus saturated and zero pixels."""
    rng = np.random.default_rng(seed)
    sig = np.array([30, 30, 25, 25, 20, 20, 20], dtype=np.float64)
    f = np.clip(np.round(128 + sig * rng.standard_normal((segments, height, width, 7))), 0, 255)
    f = f.astype(np.uint8)
    f[0, 0, :7] = np.arange(7 * 7, dtype=np.uint8).reshape(7, 7) * 5
    f[-1, -1, -3:] = 255
    f[-1, -1, :3] = 0
    return f
