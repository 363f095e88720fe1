"""Generates tests/golden/ref_*.npz from the parts of the UNMODIFIED reference that import in this container
(everything on the Stage-II path that needs chumpy / psbody / ezc3d does not):

  * moshpp.rigid_transformations (numpy, scipy, cv2 only): `rigid_landmark_transform`, `perform_rigid_adjustment`
    -- the first-frame rigid adjustment, row a9 of SURVEY.md section 8, chmosh.py:634;
  * moshpp.tools.c3d: header / parameter parsing of a c3d file written by moshpp_b200.c3d_io (row f-1).  Its frame
    reader raises OverflowError under numpy 2 (`int32 & 0x80008000`), so only the metadata is pinned.

The vectors travel with the repository; /root/reference is needed only to regenerate them:

    python tests/golden/make_reference_vectors.py
"""
import os
import sys
import tempfile
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = '/root/reference/src'
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)


def main():
    from moshpp import rigid_transformations as ref_rigid      # the reference, unmodified
    from moshpp.tools import c3d as ref_c3d
    from moshpp_b200 import c3d_io

    rng = np.random.default_rng(20240924)
    cases = []
    for k in range(12):
        m = int(rng.integers(3, 60))
        sim = rng.normal(0, 0.4, (m, 3)) + rng.normal(0, 1.0, 3)
        rv = rng.normal(0, 1.0, 3) * (3.0 if k % 4 == 0 else 0.7)        # some large rotations
        import cv2
        R = cv2.Rodrigues(rv)[0]
        obs = sim @ R.T + rng.normal(0, 1.0, 3) + rng.normal(0, 0.003, (m, 3))
        if k == 5:                      # reflection-prone: nearly planar, noisy
            sim[:, 2] *= 1e-3
            obs = sim @ R.T + rng.normal(0, 0.05, (m, 3))
        if k == 7:                      # NaN rows of the observation are replaced by the simulated ones (line 52)
            obs[1] = np.nan
        R_ref, T_ref = ref_rigid.rigid_landmark_transform(sim.T, obs.T)
        poses, trans = [np.zeros(9)], [np.zeros(3)]
        ref_rigid.perform_rigid_adjustment(poses, trans, [None], [np.where(np.isnan(obs), sim, obs)], [sim])
        cases.append(dict(sim=sim, obs=obs, R=R_ref, T=T_ref.ravel(), rv=poses[0][:3].copy(), trans=trans[0].copy()))
    np.savez_compressed(os.path.join(HERE, 'ref_rigid.npz'),
                        **{f'{k}_{i}': v for i, c in enumerate(cases) for k, v in c.items()}, n=np.array(len(cases)))
    print('ref_rigid.npz:', len(cases), 'cases')

    d = tempfile.mkdtemp(prefix='mosh_refc3d_')
    F, L = 9, 6
    data = rng.normal(0, 500, (F, L, 3))
    data[2, 1] = np.nan
    labels = ['LFHD', 'RFHD', 'C7', 'T10', '*12', 'EXTRA1']
    fn = os.path.join(d, 'written_by_c3d_io.c3d')
    c3d_io.write_c3d(fn, data, labels, frame_rate=120.0)
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        with open(fn, 'rb') as h:
            r = ref_c3d.Reader(h)
            meta = dict(point_rate=np.array(r.point_rate), point_scale=np.array(r.point_scale),
                        point_used=np.array(r.point_used), first_frame=np.array(r.first_frame),
                        last_frame=np.array(r.last_frame), labels=np.array([s.strip() for s in r.point_labels]))
    with open(fn, 'rb') as h:
        blob = np.frombuffer(h.read(), dtype=np.uint8)
    np.savez_compressed(os.path.join(HERE, 'ref_c3d.npz'), file_bytes=blob, data=data, in_labels=np.array(labels), **meta)
    print('ref_c3d.npz:', {k: (v.tolist() if v.size < 8 else v.shape) for k, v in meta.items()})


if __name__ == '__main__':
    main()
