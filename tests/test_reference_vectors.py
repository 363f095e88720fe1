"""Pins against the UNMODIFIED reference: vectors produced by the reference's own importable modules
(tests/golden/make_reference_vectors.py, run where /root/reference exists) -- `moshpp.rigid_transformations` for the
first-frame rigid adjustment and `moshpp.tools.c3d` for the c3d metadata (row f-1).  The rest
of the Stage-II path cannot be imported (chumpy / psbody / ezc3d absent) and stays pinned by the oracle's own checks."""
import os

import numpy as np

from moshpp_b200 import c3d_io
from oracle import rigid

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def test_oracle_rigid_adjustment_equals_reference():
    g = np.load(os.path.join(GOLD, 'ref_rigid.npz'))
    for i in range(int(g['n'])):
        sim, obs = g[f'sim_{i}'], g[f'obs_{i}']
        R, T = rigid.rigid_landmark_transform(sim.T, obs.T)
        assert np.abs(R - g[f'R_{i}']).max() < 1e-12 and np.abs(T.ravel() - g[f'T_{i}']).max() < 1e-12
        rv, tr = rigid.perform_rigid_adjustment(sim, np.where(np.isnan(obs), sim, obs))
        assert np.abs(tr - g[f'trans_{i}']).max() < 1e-12
        # the reference goes through cv2.Rodrigues; same rotation (axis-angle is unique below pi)
        Ra, Rb = rigid.rodrigues(rv), rigid.rodrigues(g[f'rv_{i}'])
        assert np.abs(Ra - Rb).max() < 1e-9
        assert np.abs(rv - g[f'rv_{i}']).max() < 1e-8


def test_c3d_writer_output_is_what_the_reference_parser_read(tmp_path):
    g = np.load(os.path.join(GOLD, 'ref_c3d.npz'))
    fn = str(tmp_path / 'w.c3d')
    labels = [str(s) for s in g['in_labels']]
    c3d_io.write_c3d(fn, g['data'], labels, frame_rate=120.0)
    with open(fn, 'rb') as h:
        assert np.array_equal(np.frombuffer(h.read(), dtype=np.uint8), g['file_bytes'])      # the file the reference parsed
    # what the reference's Reader made of that file
    assert float(g['point_rate']) == 120.0 and int(g['point_used']) == len(labels)
    assert int(g['last_frame']) - int(g['first_frame']) + 1 == g['data'].shape[0]
    assert [str(s) for s in g['labels']] == labels
    assert float(g['point_scale']) < 0                                                      # floating-point storage
    pts, lab, rate = c3d_io.read_c3d(fn)
    assert lab == labels and rate == 120.0
    assert np.allclose(pts, g['data'], rtol=0, atol=1e-3, equal_nan=True)                   # float32 millimetres
