"""tests/golden/generator_states.npz: keys, shapes and digests of the state_dict the
REFERENCE constructor produces for every generator choice (build container only):

    python tests/golden/make_generator_states.py

``Model(51, 3, 'mv', base_model='resnet18', arch_estimator=X, att=A, gen_flow_ds_factor=F,
use_databn=0)`` of code/dmcnet/model.py under ``torch.manual_seed(1)`` (``pretrained=True`` ->
random init, oracle/ref_loader.py).  Only ``gen_flow_model.*`` entries are stored.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_loader as R                       # noqa: E402
from oracle.digest import digest                         # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
CASES = [('ContextNetwork', 0, 0), ('ContextNetwork', 1, 0), ('ContextNetwork', 0, 4), ('ContextNetwork', 1, 4),
         ('DenseNet', 0, 0), ('DenseNetSmall', 0, 0), ('DenseNetTiny', 0, 0),
         ('DenseNetTinyEarlyFusionSum', 0, 0), ('DenseNetTinyEarlyFusionStack', 0, 0)]


def reference_state(arch, att, ds):
    torch.manual_seed(1)
    m = R.build_reference_model('dmcnet', 51, 3, 'mv', base_model='resnet18', arch_estimator=arch, att=att,
                                gen_flow_ds_factor=ds, use_databn=0, gen_flow_or_delta=1)
    return m.state_dict()


def main():
    out = {}
    for arch, att, ds in CASES:
        sd = reference_state(arch, att, ds)
        tag = '%s.att%d.ds%d' % (arch, att, ds)
        keys = [k for k in sd if k.startswith('gen_flow_model')]
        out[tag + '.keys'] = np.array(keys)
        out[tag + '.shapes'] = np.array([','.join(map(str, sd[k].shape)) for k in keys])
        out[tag + '.digests'] = np.stack([digest(sd[k].float()) for k in keys])
        out[tag + '.params'] = np.int64(sum(sd[k].numel() for k in keys if 'running' not in k and 'tracked' not in k))
        print('%-40s %3d tensors %8d params' % (tag, len(keys), int(out[tag + '.params'])))
    np.savez_compressed(os.path.join(HERE, 'generator_states.npz'), **out)


if __name__ == '__main__':
    main()
