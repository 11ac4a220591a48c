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


def resize_linear_u8(src: np.ndarray, dsize: Tuple[int, int]) -> np.ndarray:
    """``cv2.resize(src, dsize, interpolation=cv2.INTER_LINEAR)`` for uint8 ``src`` [H, W] or [H, W, C],
    dsize = (width, height).  Third-party arithmetic absent from /root/reference (OpenCV, version
    unpinned by the reference): its published algorithm is restated -- positions in double, fractions
    in float, 11-bit coefficients (cvRound), horizontal pass with the fraction reset at the borders,
    vertical pass with clamped row indices, fixed-point combine
    ((b0*(D0>>4))>>16) + ((b1*(D1>>4))>>16) + 2 >> 2 -- and pinned bit for bit against the installed
    cv2 (tests/test_extensions_cpu.py); same size returns a copy."""
    if src.ndim == 3:
        return np.stack([resize_linear_u8(src[..., c], dsize) for c in range(src.shape[2])], axis=2)
    dw, dh = dsize
    sh, sw = src.shape

    def axis(ssize, dsize_, horizontal):
        scale = 1.0 / (dsize_ / ssize)
        i0, i1, a = np.zeros(dsize_, np.int64), np.zeros(dsize_, np.int64), np.zeros((dsize_, 2), np.int32)
        for d in range(dsize_):
            f = np.float32((d + 0.5) * scale - 0.5)
            s = int(np.floor(f))
            f = np.float32(f - np.float32(s))
            if horizontal:
                if s < 0:
                    f, s = np.float32(0), 0
                if s >= ssize - 1:
                    f, s = np.float32(0), ssize - 1
            i0[d], i1[d] = min(max(s, 0), ssize - 1), min(max(s + 1, 0), ssize - 1)
            a[d, 0] = int(np.rint(np.float32((np.float32(1.0) - f) * np.float32(2048))))
            a[d, 1] = int(np.rint(np.float32(f * np.float32(2048))))
        return i0, i1, a
    x0, x1, xa = axis(sw, dw, True)
    y0, y1, ya = axis(sh, dh, False)
    s32 = src.astype(np.int32)
    rows = s32[:, x0] * xa[:, 0] + s32[:, x1] * xa[:, 1]
    d0, d1 = rows[y0], rows[y1]
    b0, b1 = ya[:, 0:1], ya[:, 1:2]
    return ((((b0 * (d0 >> 4)) >> 16) + ((b1 * (d1 >> 4)) >> 16) + 2) >> 2).astype(np.uint8)


def resize_group(img_group: Sequence[np.ndarray], size: Tuple[int, int]):
    """The resize of GroupMultiScaleCrop / GroupScale (code/dmcnet/transforms.py:60-76, :131-138):
    flow and mv channel by channel (resize_mv, :116-118), the residual as one 3-channel image."""
    return [np.concatenate((resize_linear_u8(img[:, :, :4], size), resize_linear_u8(img[:, :, 4:], size)), axis=2)
            for img in img_group]


def multi_scale_crop(img_group: Sequence[np.ndarray], crop_w: int, crop_h: int, offset_w: int, offset_h: int,
                     input_size=(224, 224)):
    """GroupMultiScaleCrop.__call__ for an already sampled crop (transforms.py:124-140); the reference
    names the ROW extent / offset 'w'."""
    crops = [img[offset_w:offset_w + crop_w, offset_h:offset_h + crop_h] for img in img_group]
    return resize_group(crops, (input_size[0], input_size[1]))


def scale_center_crop(img_group: Sequence[np.ndarray], scale_size: int = 256, crop_size: int = 224):
    """GroupScale(scale_size) then GroupCenterCrop(crop_size): transforms.py:36-44, :60-76
    (the validation / 1-crop test transform, code/dmcnet/train.py:98-101)."""
    scaled = resize_group(img_group, (scale_size, scale_size))
    h, w, _ = scaled[0].shape
    hs, ws = (h - crop_size) // 2, (w - crop_size) // 2
    return [img[hs:hs + crop_size, ws:ws + crop_size] for img in scaled]


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
