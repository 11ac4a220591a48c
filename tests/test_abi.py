"""CPU: the C-ABI library builds, loads and exports every symbol include/dmc_b200.h declares."""
import ctypes
import os

import pytest

from dmcnet_b200 import _native


@pytest.fixture(scope='module')
def lib():
    _native.build()
    return _native.lib()


def test_header_symbols_exported(lib):
    syms = _native.exported_symbols()
    assert len(syms) >= 40
    missing = [s for s in syms if not hasattr(lib, s)]
    assert not missing, missing


def test_abi_version_and_error_string(lib):
    lib.dmc_abi_version.restype = ctypes.c_int
    assert lib.dmc_abi_version() == 1
    assert isinstance(_native.last_error(), str)


def test_header_is_in_sync_with_sources():
    """tools/gen_header.py output == committed header (the header is generated from the .cu files)."""
    import subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    before = open(os.path.join(root, 'include', 'dmc_b200.h')).read()
    subprocess.check_call([sys.executable, os.path.join(root, 'tools', 'gen_header.py')])
    assert open(os.path.join(root, 'include', 'dmc_b200.h')).read() == before


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    monkeypatch.setattr(_native, '_lib', None)
    monkeypatch.setattr(_native, 'LIB_PATH', str(tmp_path / 'nope.so'))
    with pytest.raises(RuntimeError, match='no CPU/PyTorch fallback'):
        _native.lib()


def test_product_code_never_touches_the_oracle():
    """oracle/ is test infrastructure: nothing under dmcnet_b200/, examples/ or tools/ may import it
    (bench.py's CPU baseline legs and __graft_entry__.smoke() are the two sanctioned exceptions)."""
    import glob
    import os
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    offenders = []
    for pat in ('dmcnet_b200/**/*.py', 'examples/*.py', 'tools/*.py'):
        for path in glob.glob(os.path.join(root, pat), recursive=True):
            text = open(path).read()
            if re.search(r'^\s*(from|import)\s+oracle\b', text, re.M):
                offenders.append(os.path.relpath(path, root))
    assert offenders == []


def test_argument_errors_are_reported_without_touching_the_device(lib):
    """Entry points validate their arguments before any CUDA call: -1 and a message from
    dmc_last_error(), never an exception or a launch (include/dmc_b200.h conventions)."""
    vp, ci, cl, cf = ctypes.c_void_p, ctypes.c_int, ctypes.c_long, ctypes.c_float
    null, junk = vp(0), vp(4096)           # never dereferenced: every call below fails validation first
    lib.dmc_flow_loss_head.restype = ci
    assert lib.dmc_flow_loss_head(ci(7), junk, junk, cl(16), cf(1.0), null, cl(16), cl(16), junk, null) == -1
    assert 'kind' in _native.last_error()
    assert lib.dmc_flow_loss_head(ci(1), junk, junk, cl(18), cf(1.0), null, cl(18), cl(18), junk, null) == -1
    assert 'multiples of 4' in _native.last_error()
    lib.dmc_unpack_normalize_u8.restype = ci
    args = (cf(0.226), cf(0.229), cf(0.224), cf(0.225), null, junk, junk, null)
    assert lib.dmc_unpack_normalize_u8(junk, ci(1), ci(3), ci(3), *args) == -1
    assert 'multiple of 4' in _native.last_error()
    assert lib.dmc_unpack_normalize_u8(vp(4098), ci(1), ci(4), ci(4), *args) == -1
    assert 'aligned' in _native.last_error()
    lib.dmc_unpack_normalize_flip_u8.restype = ci
    assert lib.dmc_unpack_normalize_flip_u8(junk, null, ci(1), ci(4), ci(4), *args) == -1
    assert 'null pointer' in _native.last_error()
    lib.dmc_flow_block_mean_u8.restype = ci
    assert lib.dmc_flow_block_mean_u8(junk, ci(1), ci(8), ci(8), ci(0), cf(0.226), junk, null) == -1
    lib.dmc_flow_block_mean_flip_u8.restype = ci
    assert lib.dmc_flow_block_mean_flip_u8(junk, junk, ci(1), ci(8), ci(10), ci(4), cf(0.226), junk, null) == -1
    assert 'multiple of factor' in _native.last_error()
    lib.dmc_crop_resize_u8.restype = ci
    assert lib.dmc_crop_resize_u8(junk, ci(1), ci(8), ci(8), null, ci(1), junk, ci(4), ci(4), null) == -1
    assert 'null pointer' in _native.last_error()
    assert lib.dmc_crop_resize_u8(junk, ci(1), ci(8), ci(8), junk, ci(0), junk, ci(4), ci(4), null) == -1
    lib.dmc_mse_head.restype = ci
    assert lib.dmc_mse_head(junk, junk, cl(10), cf(1.0), null, cl(10), cl(10), junk, null) == -1
