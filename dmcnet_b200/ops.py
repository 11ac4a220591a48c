"""Thin torch-tensor wrappers over the C-ABI (include/dmc_b200.h).

PyTorch is used here only to own device memory and to name the CUDA stream;
every function forwards raw pointers and sizes to ``libdmc_b200.so``.  Tensors
must be CUDA, contiguous and of the stated dtype; violations raise.
"""
from __future__ import annotations

import ctypes
import os
from typing import Optional, Sequence

import torch

from . import _native

c_int, c_long, c_float, c_double, c_void_p = (ctypes.c_int, ctypes.c_long, ctypes.c_float,
                                              ctypes.c_double, ctypes.c_void_p)


def _ptr(t: Optional[torch.Tensor], dtype=None) -> c_void_p:
    if t is None:
        return c_void_p(0)
    if not t.is_cuda:
        raise RuntimeError('dmcnet_b200: expected a CUDA tensor (no CPU path exists)')
    if not t.is_contiguous():
        raise RuntimeError('dmcnet_b200: expected a contiguous tensor')
    if dtype is not None and t.dtype != dtype:
        raise RuntimeError('dmcnet_b200: expected dtype %s, got %s' % (dtype, t.dtype))
    return c_void_p(t.data_ptr())


def _stream() -> c_void_p:
    return c_void_p(torch.cuda.current_stream().cuda_stream)


def _iarr(v: Sequence[int]):
    return (c_int * len(v))(*[int(x) for x in v])


_hook = None          # profiling hook: callable(name, args) -> context manager, or None


def pad_hi() -> int:
    """High-side padding of the pixel-major layout (0: shared zero ring, see csrc/common.cuh)."""
    return int(_native.lib().dmc_layout_pad_hi())


def padded(h: int) -> int:
    """Padded extent of a pixel-major dimension of h pixels."""
    return h + 1 + pad_hi()


def set_call_hook(hook) -> None:
    global _hook
    _hook = hook


_SYNC_DEBUG = bool(os.environ.get('DMC_SYNC_DEBUG'))


def _call(name: str, *args) -> None:
    fn = getattr(_native.lib(), name)
    fn.restype = c_int
    if _SYNC_DEBUG:
        _native.check(fn(*args), name)
        try:
            torch.cuda.synchronize()
        except RuntimeError as e:
            raise RuntimeError('kernel failure inside %s(%s): %s' % (
                name, ', '.join(str(getattr(a, 'value', a)) for a in args), e)) from e
        return
    if _hook is None:
        _native.check(fn(*args), name)
    else:
        with _hook(name, args):
            _native.check(fn(*args), name)


def launch_count() -> int:
    fn = _native.lib().dmc_launch_count
    fn.restype = ctypes.c_longlong
    return int(fn())


def reset_launch_count() -> None:
    _native.lib().dmc_reset_launch_count()


F32, BF16, I64, U8, F64 = torch.float32, torch.bfloat16, torch.int64, torch.uint8, torch.float64


# ---------------------------------------------------------------- tap GEMMs
def tap_gemm(A_hi, A_lo, B_hi, B_lo, D, *, a_phases, a_rows, K, b_slices, N, M, ldD, Hp, Wp,
             shift, phase, bsel, engine='tc', stats=None, bw=None):
    """bw = (Y, act_hi, gb_or_None, mean, invstd): fused BatchNorm-backward reduction (tc engine)."""
    bY, bact, bgb, bmean, binv = bw if bw is not None else (None, None, None, None, None)
    name = 'dmc_tc_tap_gemm' if engine == 'tc' else 'dmc_simt_tap_gemm'
    _call(name, _ptr(A_hi, BF16), _ptr(A_lo, BF16), c_int(a_phases), c_long(a_rows), c_int(K),
          _ptr(B_hi, BF16), _ptr(B_lo, BF16), c_int(b_slices), c_int(N), _ptr(D, F32), c_long(M),
          c_int(ldD), c_int(Hp), c_int(Wp), c_int(len(shift)), _iarr(shift), _iarr(phase),
          _iarr(bsel), _ptr(stats, F64), _ptr(bY, F32), _ptr(bact, BF16), _ptr(bgb, F32),
          _ptr(bmean, F32), _ptr(binv, F32), _stream())


def wgrad_gemm(G_hi, G_lo, X_hi, X_lo, dW, *, P, Cout, x_phases, Cin, shift, phase, bsel,
               engine='tc', oihw_taps=0, workspace=None):
    name = 'dmc_tc_wgrad' if engine == 'tc' else 'dmc_simt_wgrad'
    _call(name, _ptr(G_hi, BF16), _ptr(G_lo, BF16), c_long(P), c_int(Cout), _ptr(X_hi, BF16),
          _ptr(X_lo, BF16), c_int(x_phases), c_int(Cin), _ptr(dW, F32), c_int(len(shift)),
          _iarr(shift), _iarr(phase), _iarr(bsel), c_int(oihw_taps), _ptr(workspace, F32),
          c_long(0 if workspace is None else workspace.numel()), _stream())


def wgrad_workspace_floats(P, Cout, Cin, ntaps):
    """Split-K workspace (floats) the tensor-core wgrad needs for this shape."""
    fn = _native.lib().dmc_tc_wgrad_workspace
    fn.restype = c_long
    return int(fn(c_long(P), c_int(Cout), c_int(Cin), c_int(ntaps)))


# ---------------------------------------------------------------- pixel-major classifier helpers
def memset_zero(t):
    _call('dmc_memset_zero', _ptr(t), c_long(t.numel() * t.element_size()), _stream())


def weight_prep(w, Cout, Cin, taps, W_hi, W_lo, Wt_hi=None, Wt_lo=None):
    _call('dmc_weight_prep', _ptr(w, F32), c_int(Cout), c_int(Cin), c_int(taps), _ptr(W_hi, BF16),
          _ptr(W_lo, BF16), _ptr(Wt_hi, BF16), _ptr(Wt_lo, BF16), _stream())


def weight_prep_multi(params, chunks, nchunks):
    _call('dmc_weight_prep_multi', _ptr(params, F32), _ptr(chunks, I64), c_int(nchunks), _stream())


def wgrad_unpack(dWs, grad, Cout, Cin, taps):
    _call('dmc_wgrad_unpack', _ptr(dWs, F32), _ptr(grad, F32), c_int(Cout), c_int(Cin), c_int(taps),
          _stream())


def bn_stats(Y, P, C, sums):
    _call('dmc_bn_stats', _ptr(Y, F32), c_long(P), c_int(C), _ptr(sums, F64), _stream())


def bn_finalize(sums, count, gamma, beta, rmean, rvar, nbt, momentum, eps, C, scale, shift, mean,
                invstd):
    _call('dmc_bn_finalize', _ptr(sums, F64), c_double(count), _ptr(gamma, F32), _ptr(beta, F32),
          _ptr(rmean, F32), _ptr(rvar, F32), _ptr(nbt, I64), c_float(momentum), c_float(eps),
          c_int(C), _ptr(scale, F32), _ptr(shift, F32), _ptr(mean, F32), _ptr(invstd, F32), _stream())


def bn_eval_coeffs(gamma, beta, rmean, rvar, eps, C, scale, shift):
    _call('dmc_bn_eval_coeffs', _ptr(gamma, F32), _ptr(beta, F32), _ptr(rmean, F32), _ptr(rvar, F32),
          c_float(eps), c_int(C), _ptr(scale, F32), _ptr(shift, F32), _stream())


def bn_apply(Y, scale, shift, P, C, Hp, Wp, relu, out_hi, out_lo, res_hi=None, res_lo=None,
             resY=None, res_scale=None, res_shift=None):
    _call('dmc_bn_apply', _ptr(Y, F32), _ptr(scale, F32), _ptr(shift, F32), c_long(P), c_int(C),
          c_int(Hp), c_int(Wp), c_int(1 if relu else 0), _ptr(res_hi, BF16), _ptr(res_lo, BF16),
          _ptr(resY, F32), _ptr(res_scale, F32), _ptr(res_shift, F32), _ptr(out_hi, BF16),
          _ptr(out_lo, BF16), _stream())


def bn_bwd_reduce(g_a, g_b, act_hi, Y, mean, invstd, P, C, Hp, Wp, sums2):
    _call('dmc_bn_bwd_reduce', _ptr(g_a, F32), _ptr(g_b, F32), _ptr(act_hi, BF16), _ptr(Y, F32),
          _ptr(mean, F32), _ptr(invstd, F32), c_long(P), c_int(C), c_int(Hp), c_int(Wp),
          _ptr(sums2, F64), _stream())


def bn_bwd_apply(g_a, g_b, act_hi, Y, mean, invstd, gamma, sums2, count, P, C, Hp, Wp, G_hi, G_lo,
                 dz_out, dgamma, dbeta):
    _call('dmc_bn_bwd_apply', _ptr(g_a, F32), _ptr(g_b, F32), _ptr(act_hi, BF16), _ptr(Y, F32),
          _ptr(mean, F32), _ptr(invstd, F32), _ptr(gamma, F32), _ptr(sums2, F64), c_double(count),
          c_long(P), c_int(C), c_int(Hp), c_int(Wp), _ptr(G_hi, BF16), _ptr(G_lo, BF16),
          _ptr(dz_out, F32), _ptr(dgamma, F32), _ptr(dbeta, F32), _stream())


def phase_split(in_hi, in_lo, frames, H, W, C, out_hi, out_lo):
    _call('dmc_phase_split', _ptr(in_hi, BF16), _ptr(in_lo, BF16), c_int(frames), c_int(H), c_int(W),
          c_int(C), _ptr(out_hi, BF16), _ptr(out_lo, BF16), _stream())


def phase_unsplit(inp, frames, H, W, C, out):
    _call('dmc_phase_unsplit', _ptr(inp, F32), c_int(frames), c_int(H), c_int(W), c_int(C),
          _ptr(out, F32), _stream())


def phase_unsplit_reduce(inp, frames, H, W, C, act_hi, Y, mean, invstd, out, sums2):
    """phase_unsplit + ReLU mask of the activation the gradient belongs to + both BatchNorm-backward
    reductions of the unit that produced it (sums2 += (sum dz, sum dz * xhat)); out = dz."""
    _call('dmc_phase_unsplit_reduce', _ptr(inp, F32), c_int(frames), c_int(H), c_int(W), c_int(C),
          _ptr(act_hi, BF16), _ptr(Y, F32), _ptr(mean, F32), _ptr(invstd, F32), _ptr(out, F32),
          _ptr(sums2, F64), _stream())


def avgpool(hi, lo, frames, Hp, Wp, C, pooled):
    _call('dmc_avgpool', _ptr(hi, BF16), _ptr(lo, BF16), c_int(frames), c_int(Hp), c_int(Wp),
          c_int(C), _ptr(pooled, F32), _stream())


def avgpool_bwd(dpooled, frames, Hp, Wp, C, dX):
    _call('dmc_avgpool_bwd', _ptr(dpooled, F32), c_int(frames), c_int(Hp), c_int(Wp), c_int(C),
          _ptr(dX, F32), _stream())


def split_planes(X, P, C, Hp, Wp, hi, lo):
    _call('dmc_split_planes', _ptr(X, F32), c_long(P), c_int(C), c_int(Hp), c_int(Wp), _ptr(hi, BF16),
          _ptr(lo, BF16), _stream())


# ---------------------------------------------------------------- planar (NCHW) small-channel layers
def conv_fwd(inp, in_ns, Cin, H, W, w, bias, Cout, ks, stride, out, out_ns, N, slope=1.0, mask=None,
             add=None, add_ns=0, accumulate=False):
    _call('dmc_conv_fwd', _ptr(inp, F32), c_long(in_ns), c_int(Cin), c_int(H), c_int(W), _ptr(w, F32),
          _ptr(bias, F32), c_int(Cout), c_int(ks), c_int(stride), _ptr(out, F32), c_long(out_ns),
          c_float(slope), _ptr(mask, F32), _ptr(add, F32), c_long(add_ns),
          c_int(1 if accumulate else 0), c_int(N), _stream())


def conv_dgrad(dY, dy_ns, Cout, w, Cin, ci_count, ks, stride, dX, dx_ns, H, W, N, accumulate=False):
    _call('dmc_conv_dgrad', _ptr(dY, F32), c_long(dy_ns), c_int(Cout), _ptr(w, F32), c_int(Cin),
          c_int(ci_count), c_int(ks), c_int(stride), _ptr(dX, F32), c_long(dx_ns), c_int(H), c_int(W),
          c_int(1 if accumulate else 0), c_int(N), _stream())


def weight_flip(w, Cout, Cin, ci_count, wT):
    _call('dmc_weight_flip', _ptr(w, F32), c_int(Cout), c_int(Cin), c_int(ci_count), _ptr(wT, F32),
          _stream())


def conv3x3_dgrad_fused(dY, dy_ns, Cy, H, W, wT, Cx, dX, dx_ns, N, accumulate=False, act_src=None,
                        act_ns=0, act_c1=0, act_slope=1.0):
    _call('dmc_conv3x3_dgrad_fused', _ptr(dY, F32), c_long(dy_ns), c_int(Cy), c_int(H), c_int(W),
          _ptr(wT, F32), c_int(Cx), _ptr(dX, F32), c_long(dx_ns), c_int(1 if accumulate else 0),
          _ptr(act_src, F32), c_long(act_ns), c_int(act_c1), c_float(act_slope), c_int(N), _stream())


def conv_wgrad(inp, in_ns, Cin, H, W, dY, dy_ns, Cout, ks, stride, dW, dbias, N):
    _call('dmc_conv_wgrad', _ptr(inp, F32), c_long(in_ns), c_int(Cin), c_int(H), c_int(W),
          _ptr(dY, F32), c_long(dy_ns), c_int(Cout), c_int(ks), c_int(stride), _ptr(dW, F32),
          _ptr(dbias, F32), c_int(N), _stream())


def act_bwd_planar(dA, da_ns, A, a_ns, mask, slope, C, HW, N, dPre, dp_ns):
    _call('dmc_act_bwd_planar', _ptr(dA, F32), c_long(da_ns), _ptr(A, F32), c_long(a_ns),
          _ptr(mask, F32), c_float(slope), c_int(C), c_long(HW), c_int(N), _ptr(dPre, F32),
          c_long(dp_ns), _stream())


def bn_stats_planar(X, x_ns, C, HW, N, sums):
    _call('dmc_bn_stats_planar', _ptr(X, F32), c_long(x_ns), c_int(C), c_long(HW), c_int(N),
          _ptr(sums, F64), _stream())


def bn_apply_planar(X, x_ns, scale, shift, C, HW, N, relu, out, o_ns):
    _call('dmc_bn_apply_planar', _ptr(X, F32), c_long(x_ns), _ptr(scale, F32), _ptr(shift, F32),
          c_int(C), c_long(HW), c_int(N), c_int(1 if relu else 0), _ptr(out, F32), c_long(o_ns),
          _stream())


def bn_bwd_reduce_planar(dZ, dz_ns, X, x_ns, mean, invstd, C, HW, N, sums2):
    _call('dmc_bn_bwd_reduce_planar', _ptr(dZ, F32), c_long(dz_ns), _ptr(X, F32), c_long(x_ns),
          _ptr(mean, F32), _ptr(invstd, F32), c_int(C), c_long(HW), c_int(N), _ptr(sums2, F64),
          _stream())


def bn_bwd_apply_planar(dZ, dz_ns, X, x_ns, mean, invstd, gamma, sums2, count, C, HW, N, dX, dx_ns,
                        dgamma, dbeta):
    _call('dmc_bn_bwd_apply_planar', _ptr(dZ, F32), c_long(dz_ns), _ptr(X, F32), c_long(x_ns),
          _ptr(mean, F32), _ptr(invstd, F32), _ptr(gamma, F32), _ptr(sums2, F64), c_double(count),
          c_int(C), c_long(HW), c_int(N), _ptr(dX, F32), c_long(dx_ns), _ptr(dgamma, F32),
          _ptr(dbeta, F32), _stream())


def copy_planar(src, s_ns, dst, d_ns, count, N):
    _call('dmc_copy_planar', _ptr(src, F32), c_long(s_ns), _ptr(dst, F32), c_long(d_ns),
          c_long(count), c_int(N), _stream())


# ---------------------------------------------------------------- stem, heads, optimizer
def stem_conv_tc_fwd(x, in_ns, H, W, w, wb, Y, y_ns, N):
    """Stem 7x7/2 conv (2 -> 64 channels) on the tensor cores; wb = 128*128 bf16 scratch."""
    _call('dmc_stem_conv_tc_fwd', _ptr(x, F32), c_long(in_ns), c_int(H), c_int(W), _ptr(w, F32),
          _ptr(wb, BF16), _ptr(Y, F32), c_long(y_ns), c_int(N), _stream())


def stem_conv_tc_wgrad(x, in_ns, H, W, dP, dp_ns, dW, ws, N):
    """dW[64][2][7][7] += stem conv weight gradient (tensor cores); ws from stem_wgrad_workspace_floats()."""
    _call('dmc_stem_conv_tc_wgrad', _ptr(x, F32), c_long(in_ns), c_int(H), c_int(W), _ptr(dP, F32),
          c_long(dp_ns), _ptr(dW, F32), _ptr(ws, F32), c_int(N), _stream())


def stem_wgrad_workspace_floats():
    fn = _native.lib().dmc_stem_conv_tc_wgrad_workspace
    fn.restype = c_long
    return int(fn())


def conv3x3_taps2(x, in_ns, Cin, H, W, w, bias, Cout, t0, out, out_ns, N, slope=1.0, mask=None):
    """3x3/1 conv using only the taps [t0, t0+2)^2 of w (space-to-depth form of a stride-2 conv)."""
    _call('dmc_conv3x3_taps2', _ptr(x, F32), c_long(in_ns), c_int(Cin), c_int(H), c_int(W), _ptr(w, F32),
          _ptr(bias, F32), c_int(Cout), c_int(t0), _ptr(out, F32), c_long(out_ns), c_float(slope),
          _ptr(mask, F32), c_int(N), _stream())


def s2d_planar(x, in_ns, C, H, W, S, s_ns, N):
    _call('dmc_s2d_planar', _ptr(x, F32), c_long(in_ns), c_int(C), c_int(H), c_int(W), _ptr(S, F32),
          c_long(s_ns), c_int(N), _stream())


def d2s_planar(S, s_ns, C, H, W, out, out_ns, N, accumulate=False):
    _call('dmc_d2s_planar', _ptr(S, F32), c_long(s_ns), c_int(C), c_int(H), c_int(W), _ptr(out, F32),
          c_long(out_ns), c_int(1 if accumulate else 0), c_int(N), _stream())


def s2_weight_map(src, dst, Cout, Cin, to_s2d):
    _call('dmc_s2_weight_map', _ptr(src, F32), _ptr(dst, F32), c_int(Cout), c_int(Cin),
          c_int(1 if to_s2d else 0), _stream())


def stem_pool_fwd(Y, scale, shift, N, C, H, W, out_hi, out_lo, idx):
    _call('dmc_stem_pool_fwd', _ptr(Y, F32), _ptr(scale, F32), _ptr(shift, F32), c_int(N), c_int(C),
          c_int(H), c_int(W), _ptr(out_hi, BF16), _ptr(out_lo, BF16), _ptr(idx, U8), _stream())


def stem_pool_bwd(g_a, g_b, idx, Y, scale, shift, N, C, H, W, dZ):
    _call('dmc_stem_pool_bwd', _ptr(g_a, F32), _ptr(g_b, F32), _ptr(idx, U8), _ptr(Y, F32),
          _ptr(scale, F32), _ptr(shift, F32), c_int(N), c_int(C), c_int(H), c_int(W), _ptr(dZ, F32),
          _stream())


def linear_fwd(x, w, b, M, K, N, out):
    _call('dmc_linear_fwd', _ptr(x, F32), _ptr(w, F32), _ptr(b, F32), c_int(M), c_int(K), c_int(N),
          _ptr(out, F32), _stream())


def linear_bwd(dy, x, w, M, K, N, dx, dw, db):
    _call('dmc_linear_bwd', _ptr(dy, F32), _ptr(x, F32), _ptr(w, F32), c_int(M), c_int(K), c_int(N),
          _ptr(dx, F32), _ptr(dw, F32), _ptr(db, F32), _stream())


def ce_head(logits, B, S, C, target, gscale, consensus, dlogits, out_stats):
    _call('dmc_ce_head', _ptr(logits, F32), c_int(B), c_int(S), c_int(C), _ptr(target, I64),
          c_float(gscale), _ptr(consensus, F32), _ptr(dlogits, F32), _ptr(out_stats, F32), _stream())


def mse_head(gen, flow, numel, gscale, dgen, loss_sum, frame_elems=None, dgen_ns=None):
    fe = frame_elems if frame_elems is not None else numel
    _call('dmc_mse_head', _ptr(gen, F32), _ptr(flow, F32), c_long(numel), c_float(gscale),
          _ptr(dgen, F32), c_long(fe), c_long(dgen_ns if dgen_ns is not None else fe),
          _ptr(loss_sum, F64), _stream())


def flow_loss_head(kind, gen, flow, numel, gscale, dgen, loss_sum, frame_elems=None, dgen_ns=None):
    """kind 0 MSELoss / 1 SmoothL1Loss / 2 L1Loss (code/dmcnet/train.py:166-172)."""
    fe = frame_elems if frame_elems is not None else numel
    _call('dmc_flow_loss_head', c_int(kind), _ptr(gen, F32), _ptr(flow, F32), c_long(numel),
          c_float(gscale), _ptr(dgen, F32), c_long(fe), c_long(dgen_ns if dgen_ns is not None else fe),
          _ptr(loss_sum, F64), _stream())


def unpack_normalize_u8(frames, N, H, W, div_motion, div_res, flow, mv, res):
    """uint8 [N][H][W][7] -> planar fp32 flow (optional) / mv / residual, normalised as
    code/dmcnet/dataset.py:251-263.  div_res = the three residual std values."""
    _call('dmc_unpack_normalize_u8', _ptr(frames, U8), c_int(N), c_int(H), c_int(W), c_float(div_motion),
          c_float(div_res[0]), c_float(div_res[1]), c_float(div_res[2]), _ptr(flow, F32), _ptr(mv, F32),
          _ptr(res, F32), _stream())


def flow_block_mean_u8(frames, N, H, W, factor, div_motion, flow):
    """Block-mean ("blocky") flow target of --flow_ds_factor (code/dmcnet/dataset.py:226-246)."""
    _call('dmc_flow_block_mean_u8', _ptr(frames, U8), c_int(N), c_int(H), c_int(W), c_int(factor),
          c_float(div_motion), _ptr(flow, F32), _stream())


def unpack_normalize_flip_u8(frames, flip, N, H, W, div_motion, div_res, flow, mv, res):
    """unpack_normalize_u8 with the random horizontal flip of code/dmcnet/transforms.py:47-58 folded
    in; flip: uint8 [N] on the device (non-zero = mirrored frame, x components 256 - v)."""
    _call('dmc_unpack_normalize_flip_u8', _ptr(frames, U8), _ptr(flip, U8), c_int(N), c_int(H), c_int(W),
          c_float(div_motion), c_float(div_res[0]), c_float(div_res[1]), c_float(div_res[2]),
          _ptr(flow, F32), _ptr(mv, F32), _ptr(res, F32), _stream())


def flow_block_mean_flip_u8(frames, flip, N, H, W, factor, div_motion, flow):
    _call('dmc_flow_block_mean_flip_u8', _ptr(frames, U8), _ptr(flip, U8), c_int(N), c_int(H), c_int(W),
          c_int(factor), c_float(div_motion), _ptr(flow, F32), _stream())


def crop_resize_u8(src, N, Hs, Ws, tab, frames_per_tab, out, Ho, Wo):
    """Crop + cv2-compatible bilinear resize of uint8 [N][Hs][Ws][7] stacks from host-built tables
    (code/dmcnet/transforms.py:122-140)."""
    _call('dmc_crop_resize_u8', _ptr(src, U8), c_int(N), c_int(Hs), c_int(Ws), _ptr(tab, torch.int32),
          c_int(frames_per_tab), _ptr(out, U8), c_int(Ho), c_int(Wo), _stream())


def dense_dgrad_weights(params, table, out):
    _call('dmc_dense_dgrad_weights', _ptr(params, F32), _iarr(table), _ptr(out, F32), _stream())


def adam_step(p, g, m, v, chunks, nchunks, hyper, step, beta1, beta2, eps, grad_scale=1.0):
    _call('dmc_adam_step', _ptr(p, F32), _ptr(g, F32), _ptr(m, F32), _ptr(v, F32),
          _ptr(chunks, torch.int32), c_int(nchunks), _ptr(hyper, F32), _ptr(step, torch.int32),
          c_float(beta1), c_float(beta2), c_float(eps), c_float(grad_scale), _stream())


# ---------------------------------------------------------------- discriminator on the tensor-core path
I32 = torch.int32


def tap_gemm_act(A_hi, A_lo, B_hi, B_lo, D, *, a_phases, a_rows, K, b_slices, N, M, ldD, Hp, Wp,
                 shift, phase, bsel, bias, mask=None, slope=0.2, stats=None):
    """tap_gemm whose epilogue applies D = mask[frame][n] * LeakyReLU(D + bias[n]) (csrc/gemm_tc.cu ActFuse)."""
    _call('dmc_tc_tap_gemm_act', _ptr(A_hi, BF16), _ptr(A_lo, BF16), c_int(a_phases), c_long(a_rows), c_int(K),
          _ptr(B_hi, BF16), _ptr(B_lo, BF16), c_int(b_slices), c_int(N), _ptr(D, F32), c_long(M),
          c_int(ldD), c_int(Hp), c_int(Wp), c_int(len(shift)), _iarr(shift), _iarr(phase), _iarr(bsel),
          _ptr(stats, F64), _ptr(bias, F32), _ptr(mask, F32), c_float(slope), _stream())


def weight_gather_prep(w, gmap, T, N, K, W_hi, W_lo, Wt_hi=None, Wt_lo=None, bias=None, bmap=None,
                       bias_exp=None):
    _call('dmc_weight_gather_prep', _ptr(w, F32), _ptr(gmap, I32), c_int(T), c_int(N), c_int(K),
          _ptr(W_hi, BF16), _ptr(W_lo, BF16), _ptr(Wt_hi, BF16), _ptr(Wt_lo, BF16), _ptr(bias, F32),
          _ptr(bmap, I32), _ptr(bias_exp, F32), _stream())


def weight_grad_gather(dWg, inv, n_w, R, dW, dbias_exp=None, binv=None, C=0, dbias=None):
    _call('dmc_weight_grad_gather', _ptr(dWg, F32), _ptr(inv, I32), c_int(n_w), c_int(R), _ptr(dW, F32),
          _ptr(dbias_exp, F64), _ptr(binv, I32), c_int(C), _ptr(dbias, F32), _stream())


def pm_bn_finalize(sums, cmap, Cp, C, count, gamma, beta, rmean, rvar, nbt, momentum, eps, scale, shift,
                   mean=None, invstd=None):
    _call('dmc_pm_bn_finalize', _ptr(sums, F64), _ptr(cmap, I32), c_int(Cp), c_int(C), c_double(count),
          _ptr(gamma, F32), _ptr(beta, F32), _ptr(rmean, F32), _ptr(rvar, F32), _ptr(nbt, I64),
          c_float(momentum), c_float(eps), _ptr(scale, F32), _ptr(shift, F32), _ptr(mean, F32),
          _ptr(invstd, F32), _stream())


def pm_bn_bwd_fold(sums2, cmap, Cp, C, count, gamma, invstd, coef, dgamma=None, dbeta=None):
    _call('dmc_pm_bn_bwd_fold', _ptr(sums2, F64), _ptr(cmap, I32), c_int(Cp), c_int(C), c_double(count),
          _ptr(gamma, F32), _ptr(invstd, F32), _ptr(coef, F32), _ptr(dgamma, F32), _ptr(dbeta, F32),
          _stream())


def pm_act_bwd(dZ, A, mean, invstd, coef, mask, slope, P, C, Hp, Wp, G_hi, G_lo, dbias_exp=None):
    _call('dmc_pm_act_bwd', _ptr(dZ, F32), _ptr(A, F32), _ptr(mean, F32), _ptr(invstd, F32),
          _ptr(coef, F32), _ptr(mask, F32), c_float(slope), c_long(P), c_int(C), c_int(Hp), c_int(Wp),
          _ptr(G_hi, BF16), _ptr(G_lo, BF16), _ptr(dbias_exp, F64), _stream())


def planar_to_s2d4(x, x_ns, H, W, M, out_hi, out_lo):
    _call('dmc_planar_to_s2d4', _ptr(x, F32), c_long(x_ns), c_int(H), c_int(W), c_int(M), _ptr(out_hi, BF16),
          _ptr(out_lo, BF16), _stream())


def s2d4_to_planar(dS, H, W, M, dX, dx_ns, accumulate=False):
    _call('dmc_s2d4_to_planar', _ptr(dS, F32), c_int(H), c_int(W), c_int(M), _ptr(dX, F32), c_long(dx_ns),
          c_int(1 if accumulate else 0), _stream())


def pm_linear_fwd(Z_hi, Z_lo, Wl, b, M, C, H, W, out):
    _call('dmc_pm_linear_fwd', _ptr(Z_hi, BF16), _ptr(Z_lo, BF16), _ptr(Wl, F32), _ptr(b, F32), c_int(M),
          c_int(C), c_int(H), c_int(W), _ptr(out, F32), _stream())


def pm_linear_bwd(dv, Z_hi, Z_lo, Wl, M, C, H, W, dZ, dWl=None, db=None):
    _call('dmc_pm_linear_bwd', _ptr(dv, F32), _ptr(Z_hi, BF16), _ptr(Z_lo, BF16), _ptr(Wl, F32), c_int(M),
          c_int(C), c_int(H), c_int(W), _ptr(dZ, F32), _ptr(dWl, F32), _ptr(db, F32), _stream())


def planar_to_pm_ring2(x, in_ns, C, H, W, N, out_hi, out_lo):
    _call('dmc_planar_to_pm_ring2', _ptr(x, F32), c_long(in_ns), c_int(C), c_int(H), c_int(W), c_int(N),
          _ptr(out_hi, BF16), _ptr(out_lo, BF16), _stream())


def s2d2_ring2_to_planar(D, ldD, H, W, N, dX, dx_ns, accumulate=False):
    _call('dmc_s2d2_ring2_to_planar', _ptr(D, F32), c_int(ldD), c_int(H), c_int(W), c_int(N), _ptr(dX, F32),
          c_long(dx_ns), c_int(1 if accumulate else 0), _stream())


# ---------------------------------------------------------------- ContextNetwork (ring-R layouts, LeakyReLU)
def tap_gemm_ring(A_hi, A_lo, B_hi, B_lo, D, *, a_rows, K, b_slices, N, M, ldD, Hp, Wp, ring, shift,
                  stats=None, bw=None, bw_slope=0.0, phase=None, bsel=None, a_phases=1):
    """tap_gemm on a layout with a `ring`-wide zero ring; bw = (Y, act_hi, gb_or_None, mean, invstd) with the
    LeakyReLU slope bw_slope in the fused BatchNorm-backward epilogue."""
    bY, bact, bgb, bmean, binv = bw if bw is not None else (None, None, None, None, None)
    phase = phase if phase is not None else [0] * len(shift)
    bsel = bsel if bsel is not None else list(range(len(shift)))
    _call('dmc_tc_tap_gemm_ring', _ptr(A_hi, BF16), _ptr(A_lo, BF16), c_int(a_phases), c_long(a_rows), c_int(K),
          _ptr(B_hi, BF16), _ptr(B_lo, BF16), c_int(b_slices), c_int(N), _ptr(D, F32), c_long(M), c_int(ldD),
          c_int(Hp), c_int(Wp), c_int(ring), c_int(len(shift)), _iarr(shift), _iarr(phase), _iarr(bsel),
          _ptr(stats, F64), _ptr(bY, F32), _ptr(bact, BF16), _ptr(bgb, F32), _ptr(bmean, F32), _ptr(binv, F32),
          c_float(bw_slope), _stream())


def bn_apply_lrelu(Y, scale, shift, P, C, Hp, Wp, ring, slope, out_hi, out_lo):
    _call('dmc_bn_apply_lrelu', _ptr(Y, F32), _ptr(scale, F32), _ptr(shift, F32), c_long(P), c_int(C), c_int(Hp),
          c_int(Wp), c_int(ring), c_float(slope), _ptr(out_hi, BF16), _ptr(out_lo, BF16), _stream())


def pm_bn_bwd_apply(dz, Y, mean, invstd, coef, P, C, Hp, Wp, ring, G_hi, G_lo):
    _call('dmc_pm_bn_bwd_apply', _ptr(dz, F32), _ptr(Y, F32), _ptr(mean, F32), _ptr(invstd, F32), _ptr(coef, F32),
          c_long(P), c_int(C), c_int(Hp), c_int(Wp), c_int(ring), _ptr(G_hi, BF16), _ptr(G_lo, BF16), _stream())


def planar_to_pm_ring(x, in_ns, C, Cp, H, W, R, N, out_hi, out_lo, out_f32=None, act_hi=None, slope=1.0):
    _call('dmc_planar_to_pm_ring', _ptr(x, F32), c_long(in_ns), c_int(C), c_int(Cp), c_int(H), c_int(W), c_int(R),
          c_int(N), _ptr(out_hi, BF16), _ptr(out_lo, BF16), _ptr(out_f32, F32), _ptr(act_hi, BF16), c_float(slope),
          _stream())


def pm_ring_to_planar(Z_hi, Z_lo, Cp, C, H, W, R, N, add, add_ns, out, out_ns):
    _call('dmc_pm_ring_to_planar', _ptr(Z_hi, BF16), _ptr(Z_lo, BF16), c_int(Cp), c_int(C), c_int(H), c_int(W),
          c_int(R), c_int(N), _ptr(add, F32), c_long(add_ns), _ptr(out, F32), c_long(out_ns), _stream())


# ---------------------------------------------------------------- gen_flow_ds_factor helpers
def avgpool_planar(x, planes, H, W, f, out):
    _call('dmc_avgpool_planar', _ptr(x, F32), c_int(planes), c_int(H), c_int(W), c_int(f), _ptr(out, F32), _stream())


def tile_repeat(x, planes, h, w, f, out):
    _call('dmc_tile_repeat', _ptr(x, F32), c_int(planes), c_int(h), c_int(w), c_int(f), _ptr(out, F32), _stream())


def tile_sum(d_out, do_ns, C, h, w, f, N, d_in, di_ns, accumulate=False):
    _call('dmc_tile_sum', _ptr(d_out, F32), c_long(do_ns), c_int(C), c_int(h), c_int(w), c_int(f), c_int(N),
          _ptr(d_in, F32), c_long(di_ns), c_int(1 if accumulate else 0), _stream())


def att_flow_loss_head(kind, gen, flow, att, numel, gscale, dgen, datt, loss_sum, frame_elems=None, dgen_ns=None):
    """criterion(att * gen, att * flow) of --att 1 (code/dmcnet/train.py:246-247); kinds as flow_loss_head."""
    fe = frame_elems if frame_elems is not None else numel
    _call('dmc_att_flow_loss_head', c_int(kind), _ptr(gen, F32), _ptr(flow, F32), _ptr(att, F32), c_long(numel),
          c_float(gscale), _ptr(dgen, F32), c_long(fe), c_long(dgen_ns if dgen_ns is not None else fe),
          _ptr(datt, F32), _ptr(loss_sum, F64), _stream())


# ---------------------------------------------------------------- inference with BatchNorm folded in
def weight_fold_prep(w, scale, shift, Cout, Cin, T, Np, Kp, W_hi, W_lo, bias_out):
    _call('dmc_weight_fold_prep', _ptr(w, F32), _ptr(scale, F32), _ptr(shift, F32), c_int(Cout), c_int(Cin),
          c_int(T), c_int(Np), c_int(Kp), _ptr(W_hi, BF16), _ptr(W_lo, BF16), _ptr(bias_out, F32), _stream())


def tap_gemm_fold(A_hi, A_lo, B_hi, B_lo, *, a_phases, a_rows, K, b_slices, N, M, Hp, Wp, shift, phase, bsel,
                  bias, slope, ring=1, D=None, res_f32=None, res_hi=None, res_lo=None, out_hi=None, out_lo=None):
    """out = LeakyReLU_slope(A * B + bias [+ residual]) -> bf16 hi/lo planes (or fp32 D); csrc/gemm_tc.cu."""
    _call('dmc_tc_tap_gemm_fold', _ptr(A_hi, BF16), _ptr(A_lo, BF16), c_int(a_phases), c_long(a_rows), c_int(K),
          _ptr(B_hi, BF16), _ptr(B_lo, BF16), c_int(b_slices), c_int(N), _ptr(D, F32), c_long(M), c_int(Hp),
          c_int(Wp), c_int(ring), c_int(len(shift)), _iarr(shift), _iarr(phase), _iarr(bsel), _ptr(bias, F32),
          c_float(slope), _ptr(res_f32, F32), _ptr(res_hi, BF16), _ptr(res_lo, BF16), _ptr(out_hi, BF16),
          _ptr(out_lo, BF16), _stream())


# ---------------------------------------------------------------- I3D (csrc/i3d.cu, column sub-range GEMMs)
def pack_hp(Hp: int, Tp: int, t_hi: int = 0) -> int:
    """Hp argument of the pixelwise / GEMM kernels for a 3-D map with padded temporal extent Tp (t_hi = 1: the
    last of the Tp frames of each clip is a zero frame as well)."""
    return int(Hp) | (int(Tp) << 16) | (int(t_hi) << 30)


def tap_gemm_ex(A_hi, A_lo, B_hi, B_lo, D, *, lda, a_rows, K, b_slices, N, M, ldD, Hp, Wp, shift, bsel,
                stats=None, stats_ld=0, bw=None, gb=None):
    """Tap GEMM on column sub-ranges (dmc_tc_tap_gemm_ex).  bw = (Y, act_hi, gb, mean, invstd) fuses the
    BatchNorm-backward reductions; gb alone adds a second gradient source to a plain result."""
    bY, bact, bgb, bmean, binv = bw if bw is not None else (None, None, gb, None, None)
    _call('dmc_tc_tap_gemm_ex', _ptr(A_hi, BF16), _ptr(A_lo, BF16), c_int(lda), c_long(a_rows), c_int(K),
          _ptr(B_hi, BF16), _ptr(B_lo, BF16), c_int(b_slices), c_int(N), _ptr(D, F32), c_long(M), c_int(ldD),
          c_int(Hp), c_int(Wp), c_int(len(shift)), _iarr(shift), _iarr(bsel), _ptr(stats, F64), c_int(stats_ld),
          _ptr(bY, F32), _ptr(bact, BF16), _ptr(bgb, F32), _ptr(bmean, F32), _ptr(binv, F32), _stream())


def wgrad_gemm_ex(G_hi, G_lo, X_hi, X_lo, dW, *, ldg, ldx, P, Cout, Cin, shift, bsel, workspace=None):
    _call('dmc_tc_wgrad_ex', _ptr(G_hi, BF16), _ptr(G_lo, BF16), c_int(ldg), c_long(P), c_int(Cout),
          _ptr(X_hi, BF16), _ptr(X_lo, BF16), c_int(ldx), c_int(Cin), _ptr(dW, F32), c_int(len(shift)),
          _iarr(shift), _iarr(bsel), _ptr(workspace, F32),
          c_long(0 if workspace is None else workspace.numel()), _stream())


def i3d_unpack(data, B, Cd, T, HW, mv, res, flow):
    _call('dmc_i3d_unpack', _ptr(data, F32), c_int(B), c_int(Cd), c_int(T), c_long(HW), _ptr(mv, F32),
          _ptr(res, F32), _ptr(flow, F32), _stream())


def i3d_stem_patches(x, x_ns, clips, T, H, W, A_hi, A_lo):
    _call('dmc_i3d_stem_patches', _ptr(x, F32), c_long(x_ns), c_int(clips), c_int(T), c_int(H), c_int(W),
          _ptr(A_hi, BF16), _ptr(A_lo, BF16), _stream())


def i3d_stem_patches_bwd(dA, clips, T, H, W, dX, dx_ns, accumulate):
    _call('dmc_i3d_stem_patches_bwd', _ptr(dA, F32), c_int(clips), c_int(T), c_int(H), c_int(W), _ptr(dX, F32),
          c_long(dx_ns), c_int(1 if accumulate else 0), _stream())


def maxpool3d_out_shape(in_thw, kernel, stride):
    out = (c_int * 3)()
    fn = _native.lib().dmc_maxpool3d_out_shape
    fn.restype = c_int
    _native.check(fn(_iarr(in_thw), _iarr(kernel), _iarr(stride), out), 'dmc_maxpool3d_out_shape')
    return tuple(int(v) for v in out)


def maxpool3d_fwd(in_hi, in_lo, clips, C, in_thw, kernel, stride, out_hi, out_lo, idx, in_t_hi=0):
    _call('dmc_maxpool3d_fwd', _ptr(in_hi, BF16), _ptr(in_lo, BF16), c_int(clips), c_int(C), _iarr(in_thw),
          c_int(in_t_hi), _iarr(kernel), _iarr(stride), _ptr(out_hi, BF16), _ptr(out_lo, BF16), _ptr(idx, U8),
          _stream())


def maxpool3d_bwd(gout, idx, clips, C, in_thw, kernel, stride, add, dX, in_t_hi=0):
    _call('dmc_maxpool3d_bwd', _ptr(gout, F32), _ptr(idx, U8), c_int(clips), c_int(C), _iarr(in_thw),
          c_int(in_t_hi), _iarr(kernel), _iarr(stride), _ptr(add, F32), _ptr(dX, F32), _stream())


def i3d_head_pool_fwd(hi, lo, clips, T5, H, W, C, pooled):
    _call('dmc_i3d_head_pool_fwd', _ptr(hi, BF16), _ptr(lo, BF16), c_int(clips), c_int(T5), c_int(H), c_int(W),
          c_int(C), _ptr(pooled, F32), _stream())


def i3d_head_pool_bwd(dpooled, clips, T5, H, W, C, dX):
    _call('dmc_i3d_head_pool_bwd', _ptr(dpooled, F32), c_int(clips), c_int(T5), c_int(H), c_int(W), c_int(C),
          _ptr(dX, F32), _stream())


def mul(a, m, out):
    _call('dmc_mul', _ptr(a, F32), _ptr(m, F32), c_long(a.numel()), _ptr(out, F32), _stream())


def add_i64(p, v=1):
    _call('dmc_add_i64', _ptr(p, I64), c_int(p.numel()), ctypes.c_longlong(v), _stream())


def sgd_nesterov_step(p, g, buf, chunks, nchunks, hyper, momentum, grad_scale=1.0):
    _call('dmc_sgd_nesterov_step', _ptr(p, F32), _ptr(g, F32), _ptr(buf, F32), _ptr(chunks), c_int(nchunks),
          _ptr(hyper, F32), c_float(momentum), c_float(grad_scale), _stream())


def axpy(y, x, a=1.0):
    _call('dmc_axpy', _ptr(y, F32), _ptr(x, F32), c_float(a), c_long(y.numel()), _stream())
