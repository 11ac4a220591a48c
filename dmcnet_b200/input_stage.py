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

from typing import Optional, Tuple

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
