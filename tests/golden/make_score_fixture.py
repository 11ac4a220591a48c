"""Cut tests/golden/score_files.npz out of the score files the reference ships
(build container only; needs /root/reference):

    python tests/golden/make_score_fixture.py

exp_my/hmdb51_coviar/{iframe,mv,residual}/split1/*_score_model_best.npz and
exp_my/hmdb51_gan/split1/mv_score_model_best.npz were written by the reference's own
test.py (code/dmcnet/test.py:181-198; the GAN file by code/dmcnet_GAN/test.py with the
extra validity column).  Every 30th video (51 of 1530) is kept, in the original object
layout, together with the fused accuracy the formula of code/dmcnet/combine.py:35-56 gives
on the subset and on the full files (the latter is the 64.05 % of SURVEY.md section 6).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import video_protocol as V                   # noqa: E402

REF = '/root/reference/exp_my'
FILES = {'iframe': 'hmdb51_coviar/iframe/split1/iframe_score_model_best.npz',
         'mv': 'hmdb51_coviar/mv/split1/mv_score_model_best.npz',
         'residual': 'hmdb51_coviar/residual/split1/residual_score_model_best.npz',
         'dmc_gan': 'hmdb51_gan/split1/mv_score_model_best.npz'}
WEIGHTS = {'iframe': 2.0, 'mv': 1.0, 'residual': 1.0, 'dmc_gan': 1.0}   # combine.py:22-29 defaults
HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    out = {}
    order = list(FILES)
    full = V.combine([os.path.join(REF, FILES[k]) for k in order], [WEIGHTS[k] for k in order])
    tmp = []
    for k in order:
        with np.load(os.path.join(REF, FILES[k]), allow_pickle=True) as z:
            keep = np.arange(0, len(z['names']), 30)
            out[k + '.scores'] = z['scores'][keep]
            out[k + '.labels'] = np.asarray(z['labels'])[keep]
            out[k + '.names'] = np.asarray(z['names'])[keep]
        p = os.path.join(HERE, '_tmp_%s.npz' % k)
        np.savez(p, scores=out[k + '.scores'], labels=out[k + '.labels'], names=out[k + '.names'])
        tmp.append(p)
    sub = V.combine(tmp, [WEIGHTS[k] for k in order])
    for p in tmp:
        os.remove(p)
    out['order'] = np.array(order)
    out['weights'] = np.array([WEIGHTS[k] for k in order])
    out['accuracy_subset'] = np.array(sub)
    out['accuracy_full'] = np.array(full)
    np.savez_compressed(os.path.join(HERE, 'score_files.npz'), **out)
    print('subset', sub, 'full', full, os.path.getsize(os.path.join(HERE, 'score_files.npz')), 'bytes')


if __name__ == '__main__':
    main()
