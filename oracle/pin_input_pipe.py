"""Pin ``oracle/input_pipe.py`` against the reference's own ``dataset.py`` and write the
golden fixture ``tests/golden/input_pipe.npz``.

TEST INFRASTRUCTURE ONLY.  Runs in the build container (needs /root/reference):

    python oracle/pin_input_pipe.py            # check + (re)write the fixture

``code/dmcnet/dataset.py`` is executed UNMODIFIED; what it cannot import here is stubbed:
  * ``coviar`` (the FFmpeg MPEG-4 extension, dataset.py:22-23): ``load`` returns synthetic
    decoded motion-vector / residual arrays (int32, as the C loader does), ``get_num_frames``
    a constant;
  * ``skimage.measure.block_reduce`` (dataset.py:26): bound to the restatement in
    ``oracle.input_pipe`` -- that one function is therefore NOT pinned by this run;
  * the pre-extracted TV-L1 flow JPEGs (dataset.py:179-180) are written losslessly (PNG data
    under the ``.jpg`` names the reference opens) into a temporary directory.
The transform is the identity (frames are produced at the crop size), ``is_train=False``
(deterministic frame indices).  The reference's ``CoviarDataSet.__getitem__`` outputs are
compared bit for bit with ``sample_from_frames`` applied to the same uint8 stacks.
"""
import importlib.util
import logging
import os
import sys
import tempfile
import types
import warnings

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import input_pipe as P                      # noqa: E402
from oracle import ref_loader as R                      # noqa: E402

GOLDEN = os.path.join(ROOT, 'tests', 'golden', 'input_pipe.npz')
CASES = [('s3_64x64_f0', 3, 64, 64, 0, False), ('s3_64x64_f16', 3, 64, 64, 16, False),
         ('s2_40x36_f16', 2, 40, 36, 16, False), ('s2_32x48_f8', 2, 32, 48, 8, False),
         # GroupRandomHorizontalFlip of the reference's transforms.py forced to flip
         ('s3_64x64_f0_flip', 3, 64, 64, 0, True), ('s3_64x64_f16_flip', 3, 64, 64, 16, True),
         ('s2_40x48_f8_flip', 2, 40, 48, 8, True)]


def _load_reference_dataset(variant: str, decoded):
    """Import code/<variant>/dataset.py with the stubs described above."""
    d = os.path.join(R.REFERENCE_ROOT, 'code', variant)
    coviar = types.ModuleType('coviar')
    coviar.get_num_frames = lambda path: 10 ** 6
    coviar.load = lambda path, gop_index, gop_pos, rep, accumulate: decoded[(gop_index, gop_pos, rep)].copy()
    skimage = types.ModuleType('skimage')
    measure = types.ModuleType('skimage.measure')
    measure.block_reduce = P.block_reduce
    skimage.measure = measure
    saved = {k: sys.modules.get(k) for k in ('coviar', 'skimage', 'skimage.measure', 'transforms')}
    saved_path = list(sys.path)
    sys.modules.update({'coviar': coviar, 'skimage': skimage, 'skimage.measure': measure})
    sys.modules.pop('transforms', None)
    sys.path.insert(0, d)
    try:
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')          # SyntaxWarning: `is not 0`
            spec = importlib.util.spec_from_file_location('_ref_%s_dataset' % variant,
                                                          os.path.join(d, 'dataset.py'))
            mod = importlib.util.module_from_spec(spec)
            level = logging.getLogger().level
            spec.loader.exec_module(mod)
            logging.getLogger().setLevel(max(level, logging.WARNING))   # dataset.py:29 sets DEBUG
    finally:
        sys.path[:] = saved_path
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    return mod


def _reference_flip(variant: str):
    """The reference's own GroupRandomHorizontalFlip (code/<variant>/transforms.py:47-58)."""
    d = os.path.join(R.REFERENCE_ROOT, 'code', variant)
    spec = importlib.util.spec_from_file_location('_ref_%s_transforms' % variant, os.path.join(d, 'transforms.py'))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def reference_sample(variant: str, segments: int, height: int, width: int, factor: int, seed: int,
                     flip: bool = False):
    """Run the reference's CoviarDataSet on synthetic decoded data.
    Returns (frames uint8 [S,H,W,7] as the reference assembled them, flow, mv, residual).
    flip: the dataset's transform is the reference's GroupRandomHorizontalFlip with its coin forced
    to 'flip' (random.random -> 0.0 for that call)."""
    from PIL import Image
    rng = np.random.default_rng(seed)
    num_frames, gop = 40, 12
    decoded, captured = {}, []
    with tempfile.TemporaryDirectory() as tmp:
        data_root, flow_root = os.path.join(tmp, 'mpeg4'), os.path.join(tmp, 'flow')
        vid = os.path.join('classA', 'v_test_01')
        os.makedirs(os.path.join(data_root, 'classA'))
        flow_dir = os.path.join(flow_root, 'classA', 'v_test_01')
        os.makedirs(flow_dir)
        for idx in range(1, num_frames + 1):
            for ax in 'xy':
                img = np.clip(np.round(128 + 30 * rng.standard_normal((height, width))), 0, 255).astype(np.uint8)
                img[0, :4], img[1, :4] = 0, 255              # extremes (0 becomes 256 under the flip)
                Image.fromarray(img, mode='L').save(os.path.join(flow_dir, 'flow_%s_%05d.jpg' % (ax, idx)),
                                                    format='PNG')
            open(os.path.join(flow_dir, 'img_%05d.jpg' % idx), 'wb').close()   # /3 in dataset.py:127
        for g in range(num_frames // gop + 1):
            for p in range(gop):
                # raw decoder output: signed, beyond +-128 in places so the clip at :200-208 is exercised
                decoded[(g, p, 1)] = np.round(40 * rng.standard_normal((height, width, 2))).astype(np.int32)
                decoded[(g, p, 1)][2, :4], decoded[(g, p, 1)][3, :4] = -200, 200      # clip to 0 / 255
                decoded[(g, p, 2)] = np.round(60 * rng.standard_normal((height, width, 3))).astype(np.int32)
        lst = os.path.join(tmp, 'list.txt')
        with open(lst, 'w') as f:
            f.write('%s.avi 0 7\n' % vid)
        mod = _load_reference_dataset(variant, decoded)

        tmod = _reference_flip(variant) if flip else None

        def transform(frames):                       # records what the reference assembled
            captured.append(np.array(frames))
            if not flip:
                return frames
            saved = tmod.random.random
            tmod.random.random = lambda: 0.0
            try:
                return tmod.GroupRandomHorizontalFlip()(frames)
            finally:
                tmod.random.random = saved
        kw = dict(mv_minmaxnorm=0) if variant == 'dmcnet_GAN' else {}
        import contextlib, io
        with contextlib.redirect_stdout(io.StringIO()):
            ds = mod.CoviarDataSet(data_root, flow_root, 'hmdb51', video_list=lst, representation='mv',
                                   new_length=1, flow_ds_factor=factor, upsample_interp=False,
                                   transform=transform, num_segments=segments, is_train=False,
                                   accumulate=True, gop=gop, **kw)
            flow, mv, res, label = ds[0]
    assert label == 7 and len(captured) == 1
    return captured[0], flow, mv, res


def pin(write: bool = False, verbose: bool = True) -> int:
    """Number of compared tensors (raises on any bit difference)."""
    store, count = {}, 0
    for variant in ('dmcnet', 'dmcnet_GAN'):
        for name, S, H, W, factor, flip in CASES:
            frames, flow, mv, res = reference_sample(variant, S, H, W, factor, seed=len(name) + S + factor,
                                                     flip=flip)
            assert frames.dtype == np.uint8 and frames.shape == (S, H, W, 7)
            if flip:
                assert (frames[..., 0] == 0).any() and (frames[..., 2] == 0).any()   # v = 0 -> 256: beyond uint8
            group = P.flip_group(list(frames)) if flip else list(frames)
            o_flow, o_mv, o_res = P.sample_from_frames(group, factor)
            for tag, a, b in (('flow', o_flow, flow), ('mv', o_mv, mv), ('res', o_res, res)):
                assert a.dtype == b.dtype == torch.float32 and a.shape == b.shape, (variant, name, tag)
                assert torch.equal(a, b), (variant, name, tag, float((a - b).abs().max()))
                count += 1
            if variant == 'dmcnet':
                store[name + '.frames'] = frames
                store[name + '.flow'], store[name + '.mv'] = flow.numpy(), mv.numpy()
                store[name + '.res'] = res.numpy()
                store[name + '.factor'] = np.int64(factor)
                store[name + '.flip'] = np.bool_(flip)
            if verbose:
                print('pinned %-11s %-14s flow/mv/res bit-identical' % (variant, name))
    if write:
        np.savez_compressed(GOLDEN, **store)
        if verbose:
            print('wrote', GOLDEN, os.path.getsize(GOLDEN), 'bytes')
    return count


if __name__ == '__main__':
    if not R.reference_available():
        sys.exit('needs /root/reference')
    pin(write=True)
