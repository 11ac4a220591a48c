"""ctypes binding of the C-ABI library ``libdmc_b200.so`` (include/dmc_b200.h).

The library is built in-tree by ``build()`` (nvcc, sm_100a only) and loaded
lazily by ``lib()``.  There is no CPU or PyTorch fallback: if the shared object
is missing or an entry point fails, a ``RuntimeError`` is raised.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
import sys
from typing import List, Optional

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB_DIR = os.path.join(HERE, '_lib')
LIB_PATH = os.path.join(LIB_DIR, 'libdmc_b200.so')

SOURCES = sorted(f for f in os.listdir(CSRC) if f.endswith('.cu'))

NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17',
              '-Xcompiler', '-fPIC']

_lib: Optional[ctypes.CDLL] = None


def _nvcc() -> str:
    for p in (os.environ.get('NVCC'), '/usr/local/cuda/bin/nvcc'):
        if p and os.path.isfile(p):
            return p
    return 'nvcc'


def _stale() -> bool:
    if not os.path.isfile(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every .cu under csrc/ for sm_100a into _lib/libdmc_b200.so."""
    if not force and not _stale():
        return LIB_PATH
    os.makedirs(LIB_DIR, exist_ok=True)
    obj_dir = os.path.join(LIB_DIR, 'obj')
    os.makedirs(obj_dir, exist_ok=True)
    nvcc = _nvcc()
    objs: List[str] = []
    procs = []
    for src in SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(obj_dir, src.replace('.cu', '.o'))
        objs.append(o)
        hdr_t = max(os.path.getmtime(os.path.join(CSRC, f)) for f in os.listdir(CSRC)
                    if f.endswith('.cuh'))
        if not force and os.path.isfile(o) and os.path.getmtime(o) > max(os.path.getmtime(s), hdr_t):
            continue
        cmd = [nvcc] + NVCC_FLAGS + ['-c', s, '-o', o]
        if verbose:
            cmd.insert(1, '-Xptxas')
            cmd.insert(2, '-v')
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError('nvcc failed on %s:\n%s' % (src, out.decode(errors='replace')))
        if verbose:
            sys.stderr.write(out.decode(errors='replace'))
    cmd = [nvcc, '-shared', '-o', LIB_PATH] + objs + ['-gencode', 'arch=compute_100a,code=sm_100a']
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    if r.returncode != 0:
        raise RuntimeError('link failed:\n%s' % r.stdout.decode(errors='replace'))
    return LIB_PATH


def lib() -> ctypes.CDLL:
    """The loaded library; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise RuntimeError(
                'dmcnet_b200: %s is missing -- run `python -c "import __graft_entry__ as g; '
                'g.build()"` (needs nvcc); there is no CPU/PyTorch fallback' % LIB_PATH)
        _lib = ctypes.CDLL(LIB_PATH)
        _lib.dmc_last_error.restype = ctypes.c_char_p
    return _lib


def last_error() -> str:
    return lib().dmc_last_error().decode(errors='replace')


def check(rc: int, what: str) -> None:
    if rc != 0:
        raise RuntimeError('dmcnet_b200 native call %s failed (%d): %s' % (what, rc, last_error()))


def exported_symbols() -> List[str]:
    """Symbols declared in include/dmc_b200.h (parsed from the header)."""
    import re
    hdr = os.path.join(os.path.dirname(HERE), 'include', 'dmc_b200.h')
    with open(hdr) as f:
        text = f.read()
    return sorted(set(re.findall(r'\b(dmc_[a-z0-9_]+)\s*\(', text)))
