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
