"""Thin torch-tensor wrappers over the C-ABI (include/dmc_b200.h).

PyTorch is used here only to own device memory and to name the CUDA stream;
every function forwards raw pointers and sizes to ``libdmc_b200.so``.  Tensors
must be CUDA, contiguous and of the stated dtype; violations raise.
"""
from __future__ import annotations

import ctypes
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


def _call(name: str, *args) -> None:
    fn = getattr(_native.lib(), name)
    fn.restype = c_int
    _native.check(fn(*args), name)


F32, BF16, I64, U8, F64 = torch.float32, torch.bfloat16, torch.int64, torch.uint8, torch.float64


# ---------------------------------------------------------------- tap GEMMs
def tap_gemm(A_hi, A_lo, B_hi, B_lo, D, *, a_phases, a_rows, K, b_slices, N, M, ldD, Hp, Wp,
             shift, phase, bsel, engine='tc'):
    name = 'dmc_tc_tap_gemm' if engine == 'tc' else 'dmc_simt_tap_gemm'
    _call(name, _ptr(A_hi, BF16), _ptr(A_lo, BF16), c_int(a_phases), c_long(a_rows), c_int(K),
          _ptr(B_hi, BF16), _ptr(B_lo, BF16), c_int(b_slices), c_int(N), _ptr(D, F32), c_long(M),
          c_int(ldD), c_int(Hp), c_int(Wp), c_int(len(shift)), _iarr(shift), _iarr(phase),
          _iarr(bsel), _stream())


def wgrad_gemm(G_hi, G_lo, X_hi, X_lo, dW, *, P, Cout, x_phases, Cin, shift, phase, bsel,
               engine='tc'):
    name = 'dmc_tc_wgrad' if engine == 'tc' else 'dmc_simt_wgrad'
    _call(name, _ptr(G_hi, BF16), _ptr(G_lo, BF16), c_long(P), c_int(Cout), _ptr(X_hi, BF16),
          _ptr(X_lo, BF16), c_int(x_phases), c_int(Cin), _ptr(dW, F32), c_int(len(shift)),
          _iarr(shift), _iarr(phase), _iarr(bsel), _stream())
