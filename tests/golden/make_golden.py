"""Generates tests/golden/*.npz: outputs of the float64 oracle (sequential Stage II) on the seeded
procedural cases of tests/conftest.py.  The reference itself cannot be run here (chumpy / psbody absent,
SURVEY.md 8(c)), so these vectors pin the oracle against drift and give the GPU tests a fixed target.

    python tests/golden/make_golden.py [C1 CF ...]
"""
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

from conftest import SMALL  # noqa: E402
from moshpp_b200 import synth  # noqa: E402
from oracle import stageii  # noqa: E402


def main():
    d = tempfile.mkdtemp(prefix='mosh_golden_')
    only = sys.argv[1:]
    for name, kw in SMALL.items():
        if only and name not in only:
            continue
        case = synth.make_case(d, name, **kw)
        out = stageii.mosh_stageii(case['mocap_fname'], case['cfg'], case['markers_latent'], case['latent_labels'],
                                   case['betas'], case['marker_meta'])
        dbg = out['stageii_debug_details']
        arrs = dict(fullpose=out['fullpose'], trans=out['trans'], pose=out['_pose_reduced'], frame_ids=dbg['frame_ids'],
                    err_data=dbg['stageii_errs']['data'], j_evals=np.array(dbg['oracle_stats']['j_evals']),
                    obs_checksum=np.array([np.nansum(case['obs']), case['vis'].sum()]),
                    markers_latent=case['markers_latent'])
        if 'dmpls' in out:
            arrs['dmpls'] = out['dmpls']
        if 'expression' in out:
            arrs['expression'] = out['expression']
        np.savez_compressed(os.path.join(HERE, f'stageii_{name}.npz'), **arrs)
        print(name, out['fullpose'].shape, 'written')


if __name__ == '__main__':
    main()
