"""Device-side input stage (first half).

``U8InputStage`` performs the tail of ``CoviarDataSet.__getitem__``
(code/dmcnet/dataset.py:215-263) on the GPU: the augmented sample stays the uint8
``[S, H, W, 7]`` stack the transforms produce (channels: flow x,y | mv x,y | residual
r,g,b; dataset.py:210), crosses PCIe at 7 B/pixel instead of 28, and two streaming
kernels (``csrc/input_pipe.cu``) split, optionally block-average the flow target
(``--flow_ds_factor``, dataset.py:226-246) and normalise it -- bit-identical to the
reference arithmetic.  ``upsample_interp=True`` (scipy ``interp1d`` upsampling,
dataset.py:236-244) has no kernel and is refused.

There is no CPU path: the stage needs a CUDA device and the built library.
"""
from __future__ import annotations

import random as _random
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import ops

INPUT_STD = (0.229, 0.224, 0.225)          # code/dmcnet/dataset.py:111-112


def normalisation_divisors() -> Tuple[float, Tuple[float, float, float]]:
    """(mean(std), std) as the fp32 values the reference divides by (dataset.py:260-263)."""
    std = torch.tensor(INPUT_STD, dtype=torch.float64).reshape(1, 3, 1, 1).float()
    return float(torch.mean(std)), tuple(float(v) for v in std.reshape(-1))


def check_stack(frames_u8: torch.Tensor, frames: int, height: int, width: int) -> None:
    if frames_u8.dtype != torch.uint8:
        raise TypeError('expected a uint8 stack [..., H, W, 7], got %s' % frames_u8.dtype)
    if frames_u8.dim() < 3 or tuple(frames_u8.shape[-3:]) != (height, width, 7):
        raise ValueError('expected trailing dimensions (%d, %d, 7), got %s'
                         % (height, width, tuple(frames_u8.shape)))
    if frames_u8.numel() != frames * height * width * 7:
        raise ValueError('expected %d frames, got %d' % (frames, frames_u8.numel() // (height * width * 7)))


class U8InputStage:
    def __init__(self, frames: int, height: int = 224, width: int = 224, *, flow_ds_factor: int = 0,
                 upsample_interp: bool = False, device: Optional[torch.device] = None):
        if upsample_interp:
            raise NotImplementedError('upsample_interp=True (scipy interp1d upsampling, dataset.py:236-244) '
                                      'has no kernel; the shipped recipes use the block repeat')
        if flow_ds_factor < 0:
            raise ValueError('flow_ds_factor must be >= 0')
        if (height * width) % 4:
            raise ValueError('height*width must be a multiple of 4')
        if not torch.cuda.is_available():
            raise RuntimeError('dmcnet_b200: a CUDA device is required (no CPU path exists)')
        self.N, self.H, self.W, self.factor = frames, height, width, flow_ds_factor
        self.device = device or torch.device('cuda', torch.cuda.current_device())
        self.div_motion, self.div_res = normalisation_divisors()
        self._stack = torch.empty(frames, height, width, 7, dtype=torch.uint8, device=self.device)

    @property
    def h2d_bytes(self) -> int:
        return self._stack.numel()

    def _flip_flags(self, flip) -> torch.Tensor:
        """Per-frame uint8 flags on the device from per-clip or per-frame booleans: the reference
        draws ONE flip decision per sample, i.e. for all segments of a clip (transforms.py:48-49)."""
        f = torch.as_tensor(flip).reshape(-1).to(torch.uint8)
        if f.numel() == 0 or self.N % f.numel():
            raise ValueError('flip has %d entries for %d frames' % (f.numel(), self.N))
        f = f.repeat_interleave(self.N // f.numel())
        if self.W % 4 or (self.factor and self.W % self.factor):
            raise ValueError('flipping on the device needs a width that is a multiple of 4 and of '
                             'flow_ds_factor (so mirrored pixel groups / blocks coincide)')
        return f.to(self.device, non_blocking=True).contiguous()

    def __call__(self, frames_u8: torch.Tensor, out_flow: Optional[torch.Tensor] = None,
                 out_mv: Optional[torch.Tensor] = None, out_res: Optional[torch.Tensor] = None,
                 flip=None):
        """frames_u8: uint8 [..., H, W, 7] (host, ideally pinned, or device) holding N frames.
        flip: optional booleans, one per clip (or per frame): the random horizontal flip of
        GroupRandomHorizontalFlip (code/dmcnet/transforms.py:47-58) applied on the device -- mirrored
        frame, x components of flow and mv become 256 - v (a value a uint8 stack cannot carry, so the
        host must NOT pre-flip).
        Returns (input_flow [N,2,H,W], input_mv [N,2,H,W], input_residual [N,3,H,W]) fp32 on the
        device, written into the given tensors when provided."""
        N, H, W = self.N, self.H, self.W
        check_stack(frames_u8, N, H, W)
        f32 = dict(dtype=torch.float32, device=self.device)
        out_flow = out_flow if out_flow is not None else torch.empty(N, 2, H, W, **f32)
        out_mv = out_mv if out_mv is not None else torch.empty(N, 2, H, W, **f32)
        out_res = out_res if out_res is not None else torch.empty(N, 3, H, W, **f32)
        for t, c in ((out_flow, 2), (out_mv, 2), (out_res, 3)):
            if t.numel() != N * c * H * W:
                raise ValueError('output tensor has %d elements, expected %d' % (t.numel(), N * c * H * W))
        self._stack.copy_(frames_u8.reshape(N, H, W, 7), non_blocking=True)
        if flip is not None:
            ff = self._flip_flags(flip)
            with_flow = out_flow if self.factor == 0 else None
            ops.unpack_normalize_flip_u8(self._stack, ff, N, H, W, self.div_motion, self.div_res, with_flow,
                                         out_mv, out_res)
            if self.factor:
                ops.flow_block_mean_flip_u8(self._stack, ff, N, H, W, self.factor, self.div_motion, out_flow)
            return out_flow, out_mv, out_res
        if self.factor == 0:
            ops.unpack_normalize_u8(self._stack, N, H, W, self.div_motion, self.div_res, out_flow, out_mv,
                                    out_res)
        else:
            ops.unpack_normalize_u8(self._stack, N, H, W, self.div_motion, self.div_res, None, out_mv, out_res)
            ops.flow_block_mean_u8(self._stack, N, H, W, self.factor, self.div_motion, out_flow)
        return out_flow, out_mv, out_res


# ------------------------------------------------------------------ crop + resize (transforms.py)
INTER_RESIZE_COEF_SCALE = 2048          # OpenCV's 11-bit fixed point of the 8-bit linear resize


def resize_tables(src_len: int, dst_len: int, horizontal: bool, offset: int = 0, frame_len: Optional[int] = None):
    """Index / coefficient tables of ``cv2.resize(..., INTER_LINEAR)`` on uint8 along one axis, the way
    OpenCV computes them (double for the position, float for the fraction, coefficients rounded to 11
    bits).  The horizontal pass resets the fraction at the borders; the vertical pass only clamps the
    two row indices.  Returns int32 arrays (i0, i1, a0, a1); indices are shifted by ``offset`` (the
    crop origin inside a frame of ``frame_len`` pixels)."""
    scale = 1.0 / (dst_len / src_len)
    i0 = np.zeros(dst_len, np.int32)
    i1 = np.zeros(dst_len, np.int32)
    a0 = np.zeros(dst_len, np.int32)
    a1 = np.zeros(dst_len, np.int32)
    for d in range(dst_len):
        f = np.float32((d + 0.5) * scale - 0.5)
        s = int(np.floor(f))
        f = np.float32(f - np.float32(s))
        if horizontal:
            if s < 0:
                f, s = np.float32(0), 0
            if s >= src_len - 1:
                f, s = np.float32(0), src_len - 1
        i0[d] = min(max(s, 0), src_len - 1) + offset
        i1[d] = min(max(s + 1, 0), src_len - 1) + offset
        a0[d] = int(np.rint(np.float32((np.float32(1.0) - f) * np.float32(INTER_RESIZE_COEF_SCALE))))
        a1[d] = int(np.rint(np.float32(f * np.float32(INTER_RESIZE_COEF_SCALE))))
    if frame_len is not None and (i0.min() < 0 or i1.max() >= frame_len):
        raise ValueError('crop [%d, %d) leaves the frame of %d pixels' % (offset, offset + src_len, frame_len))
    return i0, i1, a0, a1


def crop_tables(row0: int, col0: int, rows: int, cols: int, out_h: int, out_w: int,
                frame_h: Optional[int] = None, frame_w: Optional[int] = None) -> np.ndarray:
    """One table row of ``dmc_crop_resize_u8``: ``img[row0:row0+rows, col0:col0+cols]`` resized to
    (out_h, out_w).  Layout {x0, x1, a0, a1 [out_w] | y0, y1, b0, b1 [out_h]}."""
    x = resize_tables(cols, out_w, True, col0, frame_w)
    y = resize_tables(rows, out_h, False, row0, frame_h)
    return np.concatenate(list(x) + list(y)).astype(np.int32)


def scaled_crop_tables(frame_h: int, frame_w: int, scale_h: int, scale_w: int, row0: int, col0: int,
                       out_h: int, out_w: int) -> np.ndarray:
    """GroupScale to (scale_h, scale_w) followed by the crop window [row0:row0+out_h, col0:col0+out_w]
    of the scaled frame (GroupCenterCrop / one GroupOverSample window, transforms.py:36-44, :78-114):
    the window's rows of the full-frame resize tables."""
    x = [t[col0:col0 + out_w] for t in resize_tables(frame_w, scale_w, True)]
    y = [t[row0:row0 + out_h] for t in resize_tables(frame_h, scale_h, False)]
    if len(x[0]) != out_w or len(y[0]) != out_h:
        raise ValueError('crop window leaves the scaled frame')
    return np.concatenate(x + y).astype(np.int32)


def multi_scale_crop_pairs(im_rows: int, im_cols: int, input_size: Sequence[int], scales: Sequence[float],
                           max_distort: int = 1) -> List[Tuple[int, int]]:
    """Candidate (crop_rows, crop_cols) of GroupMultiScaleCrop._sample_crop_size
    (code/dmcnet/transforms.py:142-157; the reference calls the row extent 'w')."""
    base = min(im_rows, im_cols)
    sizes = [int(base * x) for x in scales]
    crop_h = [input_size[1] if abs(x - input_size[1]) < 3 else x for x in sizes]
    crop_w = [input_size[0] if abs(x - input_size[0]) < 3 else x for x in sizes]
    return [(w, h) for i, h in enumerate(crop_h) for j, w in enumerate(crop_w) if abs(i - j) <= max_distort]


def sample_multi_scale_crop(im_rows: int, im_cols: int, input_size: Sequence[int] = (224, 224),
                            scales: Sequence[float] = (1, .875, .75), max_distort: int = 1,
                            rng=_random) -> Tuple[int, int, int, int]:
    """(row0, col0, rows, cols) drawn with the reference's calls to ``random`` in the reference's order
    (``choice`` of the pair, then ``randint`` for each offset; transforms.py:159-163, fix_crop=False)."""
    rows, cols = rng.choice(multi_scale_crop_pairs(im_rows, im_cols, list(input_size), list(scales), max_distort))
    row0 = rng.randint(0, im_rows - rows)
    col0 = rng.randint(0, im_cols - cols)
    return row0, col0, rows, cols


def oversample_offsets(scaled_rows: int, scaled_cols: int, crop_rows: int, crop_cols: int) -> List[Tuple[int, int]]:
    """The five (row0, col0) windows of GroupOverSample (fill_fix_offset(False, ...),
    transforms.py:171-181): corners and centre."""
    rs, cs = (scaled_rows - crop_rows) // 4, (scaled_cols - crop_cols) // 4
    return [(0, 0), (4 * rs, 0), (0, 4 * cs), (4 * rs, 4 * cs), (2 * rs, 2 * cs)]


class CropResizeStage:
    """Crop + resize of decoded uint8 stacks on the device: ``src`` [N, Hs, Ws, 7] -> [N, out_h, out_w, 7],
    one crop per clip, bit-identical to numpy slicing + ``cv2.resize(INTER_LINEAR)`` (installed
    OpenCV 4.x semantics).  Feed the result to ``U8InputStage`` (flip, split, normalise)."""

    def __init__(self, frames: int, src_h: int, src_w: int, out_h: int = 224, out_w: int = 224,
                 device: Optional[torch.device] = None):
        if not torch.cuda.is_available():
            raise RuntimeError('dmcnet_b200: a CUDA device is required (no CPU path exists)')
        self.N, self.Hs, self.Ws, self.Ho, self.Wo = frames, src_h, src_w, out_h, out_w
        self.device = device or torch.device('cuda', torch.cuda.current_device())
        self._src = torch.empty(frames, src_h, src_w, 7, dtype=torch.uint8, device=self.device)
        self.out = torch.empty(frames, out_h, out_w, 7, dtype=torch.uint8, device=self.device)

    def __call__(self, src_u8: torch.Tensor, tables: np.ndarray) -> torch.Tensor:
        """tables: int32 [clips, 4*out_w + 4*out_h] from ``crop_tables`` / ``scaled_crop_tables``."""
        check_stack(src_u8, self.N, self.Hs, self.Ws)
        tab = np.ascontiguousarray(np.asarray(tables, dtype=np.int32).reshape(-1, 4 * (self.Wo + self.Ho)))
        if self.N % tab.shape[0]:
            raise ValueError('%d tables for %d frames' % (tab.shape[0], self.N))
        xs, ys = tab[:, :2 * self.Wo], tab[:, 4 * self.Wo:4 * self.Wo + 2 * self.Ho]
        if xs.min() < 0 or xs.max() >= self.Ws or ys.min() < 0 or ys.max() >= self.Hs:
            raise ValueError('table indices leave the %dx%d source frame' % (self.Hs, self.Ws))
        self._src.copy_(src_u8.reshape(self.N, self.Hs, self.Ws, 7), non_blocking=True)
        dtab = torch.from_numpy(tab).to(self.device)
        ops.crop_resize_u8(self._src, self.N, self.Hs, self.Ws, dtab, self.N // tab.shape[0], self.out,
                           self.Ho, self.Wo)
        return self.out
