"""CPU suite: the device source (single-thread host build, tests/emu) against the float64 oracle.

This validates index math and control flow of the exact code nvcc compiles; the numbers that count are
the `-m gpu` twins in test_gpu_parity.py, which call the CUDA library through the C-ABI.
"""
import numpy as np
import pytest

from conftest import run_oracle
from moshpp_b200 import lib


@pytest.mark.parametrize('name', ['C1', 'C2', 'C3', 'C4', 'CF'])
def test_f64_device_source_equals_oracle(cases, emu, name):
    case = cases(name)
    out = run_oracle(case)
    res = emu(case, precision=lib.MOSH2_F64)
    dbg = out['stageii_debug_details']
    fid = dbg['frame_ids']
    assert np.array_equal(np.nonzero(res.status & lib.ST_SOLVED)[0], fid)
    assert np.abs(res.pose[fid] - out['_pose_reduced']).max() < 1e-9
    assert np.abs(res.fullpose[fid] - out['fullpose']).max() < 1e-9
    assert np.abs(res.trans[fid] - out['trans']).max() < 1e-10
    if 'dmpls' in out:
        assert np.abs(res.dmpls[fid, :out['dmpls'].shape[1]] - out['dmpls']).max() < 1e-9
    if 'expression' in out:                      # the expression coefficients are the tail of the linear block
        pk = case['pack']
        assert pk.n_expr > 0 and np.abs(out['expression'][:, :pk.n_expr]).max() > 1e-2
        assert np.abs(res.dmpls[fid, pk.n_dmpl - pk.n_expr:pk.n_dmpl] - out['expression'][:, :pk.n_expr]).max() < 1e-9
    # identical dog-leg trajectories: same number of Jacobian builds and minimisations
    assert res.counters[fid, 2].sum() == dbg['oracle_stats']['j_evals']
    assert res.counters[fid, 3].sum() == dbg['oracle_stats']['minimizations']
    for col, k in enumerate(lib.ERR_NAMES):
        if k in dbg['stageii_errs'] and k not in ('velo', 'extrap_dmpl'):
            assert np.allclose(res.errs[fid, col], dbg['stageii_errs'][k], rtol=1e-8, atol=1e-12)
    n_velo = int(((res.status[fid] & lib.ST_HAS_VELO) != 0).sum())
    assert n_velo == len(dbg['stageii_errs'].get('velo', []))


def test_f32_device_source_within_tolerance(cases, emu):
    case = cases('C2')
    out = run_oracle(case)
    res = emu(case, precision=lib.MOSH2_F32)
    fid = out['stageii_debug_details']['frame_ids']
    bd = case['pack'].body_dof
    dp = np.abs(res.pose[fid] - out['_pose_reduced'])
    assert dp[:, :bd].max() < 1e-3          # body pose, rad
    assert dp.max() < 5e-3                   # weakly observed finger PCA coefficients
    assert np.abs(res.trans[fid] - out['trans']).max() < 1e-4
    sse = out['stageii_debug_details']['stageii_errs']['data']
    assert np.abs(res.errs[fid, 0] / sse - 1).max() < 1e-2


def test_chunked_schedule_matches_oracle_chunked(cases, emu):
    case = cases('C2')
    out = run_oracle(case, chunk=(5, 2))
    res = emu(case, chunk_len=5, warmup=2)
    fid = out['stageii_debug_details']['frame_ids']
    assert np.abs(res.pose[fid] - out['_pose_reduced']).max() < 1e-9


def test_chunk_warmup_converges_to_sequential(cases, emu):
    """The reference recursion is contractive: chunks started W frames early converge to the single
    sequential pass geometrically in W."""
    case = cases('C2', frames=160)
    seq = emu(case)
    ok = (seq.status & lib.ST_SOLVED) != 0
    errs = []
    for W in (0, 16, 48, 96):
        ch = emu(case, chunk_len=32, warmup=W)
        errs.append(np.abs(ch.pose - seq.pose)[ok].max())
    assert errs[1] < 0.5 * errs[0] and errs[2] < 1e-3 and errs[3] < 1e-6, errs


def test_frames_without_markers_are_skipped(cases, emu):
    case = cases('C1')
    from conftest import dense_obs
    obs, vis = dense_obs(case)
    vis = vis.copy()
    vis[3] = False
    vis[7] = False
    res = emu(case, obs_vis=(obs, vis))
    assert res.status[3] == lib.ST_SKIPPED and res.status[7] == lib.ST_SKIPPED
    assert (res.status[[0, 1, 2, 4]] & lib.ST_SOLVED).all()
    # velocity term needs two processed predecessors (chmosh.py:624-626,656-657)
    assert not (res.status[0] & lib.ST_HAS_VELO) and not (res.status[1] & lib.ST_HAS_VELO)
    assert res.status[2] & lib.ST_HAS_VELO and res.status[4] & lib.ST_HAS_VELO


def test_light_warmup_schedule_matches_oracle_emulation(cases, emu):
    """Chunks with a light warm-up (one Step-2 iteration per frame) followed by fully solved warm-up frames: the device
    source against the oracle's independent emulation of the same schedule, incl. frames without markers inside the
    warm-up window (the warm-up is counted in solved frames)."""
    from conftest import dense_obs
    case = cases('C2')
    obs, vis = dense_obs(case)
    vis = vis.copy()
    vis[6] = False
    vis[7] = False
    from oracle import stageii
    from moshpp_b200.mocap_interface import MocapSession
    mocap = MocapSession(case['mocap_fname'], case['cfg'].mocap.unit)
    mocap.markers[6:8] = 0.0
    out = stageii.mosh_stageii(case['mocap_fname'], case['cfg'], case['markers_latent'], case['latent_labels'],
                               case['betas'], case['marker_meta'], chunk=(4, 5, 2), mocap=mocap)
    res = emu(case, chunk_len=4, warmup=5, warmup_full=2, obs_vis=(obs, vis))
    fid = out['stageii_debug_details']['frame_ids']
    assert 6 not in fid and res.status[6] == lib.ST_SKIPPED
    assert np.abs(res.pose[fid] - out['_pose_reduced']).max() < 1e-9
    # not the same numbers as with a fully solved warm-up
    full = emu(case, chunk_len=4, warmup=5, obs_vis=(obs, vis))
    assert np.abs(full.pose[fid] - res.pose[fid]).max() > 1e-6
