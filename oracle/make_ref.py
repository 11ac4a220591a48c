"""Recipe for ``oracle/_ref``: the reference's OWN implementation of the hot path, staged so that it
travels to the GPU box (``/root/reference`` exists only in the build container).

TEST / BASELINE INFRASTRUCTURE ONLY.  The reference is pure Python for this path, so
"building" it is byte-compiling the four modules the path imports, from the sources where they lie:

    code/dmcnet/model.py      code/dmcnet/transforms.py
    code/dmcnet_GAN/model.py  code/dmcnet_GAN/transforms.py

into sourceless bytecode ``oracle/_ref/code/<variant>/<module>.bc`` (a .pyc under another suffix:
snapshots drop ``*.pyc``; same relative layout, so
``oracle/ref_loader.py`` loads either root; the GPU box runs the same image, hence the same CPython
bytecode version).  No reference source text is copied anywhere; ``oracle/_ref/`` is git-ignored but not
gpurun-ignored, like the product's own built ``.so``.  ``__graft_entry__.build()`` runs this when ``/root/reference`` is present.  It is what
``bench.py --impl reference`` and the ``cpu_baseline`` leg time (``kind: "reference"``); the restated
step around it (the reference's train.py does not parse on Python >= 3.7) is oracle/ref_step.py.

    python oracle/make_ref.py
"""
import hashlib
import json
import os
import py_compile
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC_ROOT = '/root/reference'
DST_ROOT = os.path.join(HERE, '_ref')
FILES = [('code', v, f) for v in ('dmcnet', 'dmcnet_GAN') for f in ('model.py', 'transforms.py')]


def make(verbose: bool = False) -> bool:
    if not os.path.isfile(os.path.join(SRC_ROOT, *FILES[0])):
        return False
    manifest = {}
    for parts in FILES:
        src, dst = os.path.join(SRC_ROOT, *parts), os.path.join(DST_ROOT, *parts)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        dst = dst[:-3] + '.bc'
        py_compile.compile(src, cfile=dst, doraise=True, optimize=0)
        with open(dst, 'rb') as f:
            manifest['/'.join(parts)] = hashlib.sha256(f.read()).hexdigest()
    with open(os.path.join(DST_ROOT, 'MANIFEST.json'), 'w') as f:
        json.dump({'source': SRC_ROOT, 'sha256': manifest}, f, indent=1)
    if verbose:
        print('oracle/_ref: %d modules byte-compiled from %s' % (len(FILES), SRC_ROOT))
    return True


if __name__ == '__main__':
    sys.exit(0 if make(verbose=True) else 1)
