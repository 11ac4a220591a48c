"""Import the *reference's own* ``model.py`` from /root/reference (build container) or from the
sourceless byte-compiled copy staged under oracle/_ref by oracle/make_ref.py (the GPU box has no /root/reference).
TEST / BASELINE INFRASTRUCTURE ONLY.

Used by ``oracle/pin_against_reference.py`` and ``tests/golden/make_golden.py``.
Both reference variants define top-level modules ``model`` and ``transforms``
(code/dmcnet/model.py:12, code/dmcnet_GAN/model.py:12), so they are loaded one
at a time under private names.  The only patch: ``torchvision.models.resnet18(
pretrained=True)`` (code/dmcnet/model.py:305) would download weights -> it is
wrapped to ``weights=None``.
"""
import contextlib
import importlib.util
import io
import os
import sys
import warnings

_STAGED = os.path.join(os.path.dirname(os.path.abspath(__file__)), '_ref')       # oracle/make_ref.py


def _pick_root() -> str:
    """/root/reference in the build container, else the copy staged by oracle/make_ref.py."""
    env = os.environ.get('DMC_REFERENCE_ROOT')
    if env:
        return env
    if os.path.isfile(os.path.join('/root/reference', 'code', 'dmcnet', 'model.py')):
        return '/root/reference'
    return _STAGED


REFERENCE_ROOT = _pick_root()


def reference_available() -> bool:
    d = os.path.join(REFERENCE_ROOT, 'code', 'dmcnet')
    return os.path.isfile(os.path.join(d, 'model.py')) or os.path.isfile(os.path.join(d, 'model.bc'))


def load_reference_model_module(variant: str):
    """variant in {'dmcnet', 'dmcnet_GAN'} -> the imported reference module."""
    import torchvision
    d = os.path.join(REFERENCE_ROOT, 'code', variant)
    name = '_ref_%s_model' % variant
    if name in sys.modules:
        return sys.modules[name]
    saved_path = list(sys.path)
    saved_tf = sys.modules.pop('transforms', None)
    sys.path.insert(0, d)
    try:
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')          # SyntaxWarning: `is 'ContextNetwork'`
            path = os.path.join(d, 'model.py')
            if os.path.isfile(path):
                spec = importlib.util.spec_from_file_location(name, path)
            else:
                # staged, sourceless bytecode (oracle/make_ref.py): model.bc imports `transforms`,
                # which is loaded the same way first
                from importlib.machinery import SourcelessFileLoader
                tf_loader = SourcelessFileLoader('transforms', os.path.join(d, 'transforms.bc'))
                tf_spec = importlib.util.spec_from_loader('transforms', tf_loader)
                tf = importlib.util.module_from_spec(tf_spec)
                tf_loader.exec_module(tf)
                sys.modules['transforms'] = tf
                loader = SourcelessFileLoader(name, os.path.join(d, 'model.bc'))
                spec = importlib.util.spec_from_loader(name, loader)
            mod = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(mod)
    finally:
        sys.path[:] = saved_path
        sys.modules.pop('transforms', None)
        if saved_tf is not None:
            sys.modules['transforms'] = saved_tf
    sys.modules[name] = mod

    orig = {}

    class _TV:                                         # proxy seen by the reference as `torchvision`
        def __getattr__(self, k):
            return getattr(torchvision, k)

    class _Models:
        def __getattr__(self, k):
            fn = getattr(torchvision.models, k)
            if k.startswith('resnet'):
                return lambda pretrained=False, **kw: fn(weights=None, **kw)
            return fn

    tv = _TV()
    tv.__dict__['models'] = _Models()
    mod.torchvision = tv
    return mod


def build_reference_model(variant: str, *args, **kwargs):
    mod = load_reference_model_module(variant)
    with contextlib.redirect_stdout(io.StringIO()):    # the ctor prints a banner
        return mod.Model(*args, **kwargs)
