"""Host-side adapters: c3d round trip, MocapSession visibility / unit rules, synthetic forward vs oracle."""
import os
import pickle

import numpy as np

from conftest import dense_obs
from moshpp_b200 import c3d_io, pack
from moshpp_b200.mocap_interface import MocapSession


def test_c3d_round_trip(tmp_path):
    rng = np.random.default_rng(0)
    x = 1000 * rng.standard_normal((7, 11, 3))
    x[2, 3] = np.nan
    labels = [f'M{i}' for i in range(10)] + ['LONGLABEL_12345']
    fn = str(tmp_path / 't.c3d')
    c3d_io.write_c3d(fn, x, labels, frame_rate=100.0)
    p, l, r = c3d_io.read_c3d(fn)
    assert l == labels and r == 100.0
    assert np.isnan(p[2, 3]).all() and np.nanmax(np.abs(p - x)) < 1e-3


def test_mocap_session_rules(tmp_path):
    """mm -> m, '*' labels dropped, subject prefix stripped, (0,0,0) and NaN count as missing
    (tools/mocap_interface.py:186,201-212,223-225,277)."""
    mk = np.ones((3, 4, 3)) * 1000.0
    mk[0, 1] = 0.0
    mk[1, 2] = np.nan
    fn = str(tmp_path / 'm.npz')
    np.savez(fn, markers=mk, labels=np.array(['subj:A', 'B ', '*3', 'C']), frame_rate=60.0)
    s = MocapSession(fn, 'mm')
    assert s.labels == ['A', 'B', 'C'] and s.frame_rate == 60.0
    assert np.allclose(s.markers[0, 0], 1.0)
    obs, vis = s.frames_for_labels(['C', 'B', 'Z', 'A'], range(3))
    assert vis.tolist() == [[True, False, False, True], [True, True, False, True], [True, True, False, True]]
    d = s.markers_asdict()
    assert list(d[0].keys()) == ['A', 'C'] and len(d[1]) == 3
    with open(str(tmp_path / 'm.pkl'), 'wb') as f:
        pickle.dump({'markers': mk, 'labels': ['A', 'B', '*3', 'C'], 'frame_rate': 30.}, f)
    assert MocapSession(str(tmp_path / 'm.pkl'), 'm').markers[0, 0, 0] == 1000.0


def test_c1_is_a_c3d_case(cases):
    case = cases('C1')
    assert case['mocap_fname'].endswith('.c3d')
    obs, vis = dense_obs(case)
    ok = case['vis']
    assert np.array_equal(vis, ok)
    assert np.abs(obs[vis] - case['obs'][ok]).max() < 1e-6          # float32 c3d frames, mm


def test_synth_forward_equals_oracle(cases):
    from oracle import stageii
    for name in ('C2', 'C3', 'C4'):
        case = cases(name)
        sol = stageii.StageIISolver(case['cfg'], case['markers_latent'], case['latent_labels'], case['betas'], case['marker_meta'])
        for t in (0, 5):
            sol.pose[:] = case['gt_pose'][t]
            sol.trans[:] = case['gt_trans'][t]
            if sol.nd:
                sol.betas[sol.dmpl_ids] = case['gt_dmpl'][t]
            assert np.abs(sol.evaluate(False)['markers'] - case['gt_markers'][t]).max() < 1e-12
        # product attachment == oracle attachment (transformed_lm.py:59-113)
        assert np.array_equal(case['pack'].closest, sol.tc.closest[:, :3])
        assert np.abs(case['pack'].coefs - sol.tc.coefs).max() < 1e-12


def test_pack_layout(cases):
    pk = cases('C2')['pack']
    assert (pk.n_joints, pk.p_red, len(pk.free_step1), len(pk.free_step2)) == (52, 114, 63, 111)
    assert pk.pd.shape == (51, 9 * 53, 9) and pk.prior_d == 63 and pk.prior_off == 3
    assert (pk.finger_lo, pk.finger_hi) == (66, 114)
    assert not set(range(33, 39)) & set(pk.free_step2.tolist())      # toes (pose 30:36) stay frozen
    pk3 = cases('C3')['pack']
    assert len(pk3.free_step2) == 119 and pk3.n_dmpl == 8
    assert not set(range(3 + 66, 3 + 75)) & set(pk3.free_step2.tolist())   # jaw / eyes never optimised here
    pk4 = cases('C4')['pack']
    assert len(pk4.free_step1) == 6 and len(pk4.free_step2) == 30 and pk4.prior_k == 0
    parts = pack.pose_partitions('smpl', 72, False, False, False)
    assert len(parts['step1']) == 66 and parts['finger'] == []
