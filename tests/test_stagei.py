"""Stage I (reference chmosh.py:83-455): the product's block solve on per-frame device linearisations
against the oracle's dense restatement.  CPU: the device source through its host build (tests/emu); `-m gpu`: the CUDA library.
The oracle itself is checked against finite differences (its parity with chumpy is unpinned, oracle/stagei.py)."""
import numpy as np
import pytest

from conftest import EmuStageIBackend, stagei_case
from moshpp_b200 import stagei as product
from oracle import stagei as oracle


def _compare(out, ref, tol):
    assert np.abs(out['betas'] - ref['betas']).max() < tol
    assert np.abs(out['markers_latent'] - ref['markers_latent']).max() < tol
    do, dr = out['stagei_debug_details'], ref['stagei_debug_details']
    for a, b in zip(do['opt_models_pose'], dr['opt_models_pose']):
        assert np.abs(a - b).max() < tol
    for a, b in zip(do['opt_models_trans'], dr['opt_models_trans']):
        assert np.abs(a - b).max() < tol
    assert set(do['stagei_errs'].keys()) == set(dr['stagei_errs'].keys())
    for k, v in dr['stagei_errs'].items():
        assert abs(do['stagei_errs'][k] - v) <= 1e-5 * abs(v) + 100 * tol, k
    assert out['latent_labels'] == ref['latent_labels'] and out['markers_latent_vids'] == ref['markers_latent_vids']
    assert do['stagei_labels_obs'] == dr['stagei_labels_obs']
    for a, b in zip(do['stagei_markers_sim'], dr['stagei_markers_sim']):
        assert a.shape == b.shape and np.abs(a - b).max() < 10 * tol


def test_oracle_jacobian_equals_finite_differences(cases):
    case, cfg, frames = stagei_case(cases, 'C2', 3, frames=40, n_verts=1500, dropout=0.0)
    s = oracle.StageISolver(frames, cfg, case['marker_meta'])
    s.rigid_adjust()
    wts = s.weights_for(0.5)
    pose_ids = s.pose_ids_for(True)
    rng = np.random.default_rng(0)
    x0 = s.get_x(pose_ids, True)
    # (poses, translations and shape are moved off the start; the latent markers stay at their regular start positions --
    # pushed around at random some end up beside an open border of the synthetic surface, where the SIGN of the distance flips)
    ids = np.arange(len(x0))
    x0 = x0 + rng.normal(0, 0.02, x0.shape) * (ids >= s.nb + 3 * s.n_markers) + rng.normal(0, 0.3, x0.shape) * (ids < s.nb)
    r, J = s.residual(x0, True, pose_ids, True, wts, True)
    nb, M = s.nb, s.n_markers
    cols = [0, 1, nb - 1, nb, nb + 4, nb + 3 * M - 1, nb + 3 * M, nb + 3 * M + 4, nb + 3 * M + 40, len(x0) - 1]     # betas, latent markers, trans, pose
    for c in cols:
        h = 1e-6
        xp, xm = x0.copy(), x0.copy()
        xp[c] += h
        xm[c] -= h
        fd = (s.residual(xp, False, pose_ids, True, wts, True) - s.residual(xm, False, pose_ids, True, wts, True)) / (2 * h)
        assert np.abs(fd - J[:, c]).max() < 1e-5 * (np.abs(J[:, c]).max() + 1e-9), c


def test_block_solve_on_device_source_equals_oracle(cases):
    """Shape, latent markers, poses and translations of four frames: SMPL-H with finger markers and the mean hand pose in the
    canonical body (so the canonical mesh is not the template), all four annealing steps."""
    case, cfg, frames = stagei_case(cases, 'C2', 4, frames=40, n_verts=1500, dropout=0.02)
    cfg.opt_settings.maxiter = 6          # (per minimisation; the run to convergence is the -m gpu twin below)
    ref = oracle.mosh_stagei(frames, cfg, marker_meta=case['marker_meta'])
    out = product.mosh_stagei(frames, cfg, marker_meta=case['marker_meta'], backend=EmuStageIBackend())
    _compare(out, ref, 1e-9)
    st, rs = out['stagei_debug_details']['b200'], ref['stagei_debug_details']['oracle_stats']
    assert st['linearisations'] == rs['j_evals'] and st['iterations'] == rs['iterations'] and st['minimisations'] == 4
    e = ref['stagei_debug_details']['stagei_errs']
    assert e['data'] > 0 and e['surf'] > 0 and e['beta'] > 0 and e['poseH'] >= 0 and 'init_body' in e


def test_given_betas_are_kept(cases, tmp_path):
    """optimize_betas off with a betas file (chmosh.py:92-97,169-172): the shape stays, latent markers and poses are estimated."""
    case, cfg, frames = stagei_case(cases, 'C1', 3)
    cfg.moshpp.optimize_betas = False
    fn = str(tmp_path / 'betas.npz')
    np.savez(fn, betas=case['betas'])
    ref = oracle.mosh_stagei(frames, cfg, betas_fname=fn, marker_meta=case['marker_meta'])
    out = product.mosh_stagei(frames, cfg, betas_fname=fn, marker_meta=case['marker_meta'], backend=EmuStageIBackend())
    _compare(out, ref, 1e-9)
    nb = cfg.surface_model.num_betas
    assert np.array_equal(out['betas'][:nb], case['betas'][:nb]) and 'beta' not in out['stagei_debug_details']['stagei_errs']


@pytest.mark.gpu
def test_stagei_on_the_gpu_equals_oracle(cases):
    """The CUDA path: per-frame linearisations from mosh2_job_linearize, closest points and distances from mosh2_mesh_distance
    (float32 search, float64 closed forms)."""
    import time
    case, cfg, frames = stagei_case(cases, 'C2', 4, frames=40, n_verts=1500, dropout=0.02)
    cfg.opt_settings.maxiter = 12         # (48 dog-leg iterations in all; the run to convergence at BASELINE size: tools/gpu_stagei.py)
    ref = oracle.mosh_stagei(frames, cfg, marker_meta=case['marker_meta'])
    t0 = time.perf_counter()
    out = product.mosh_stagei(frames, cfg, marker_meta=case['marker_meta'])
    dt = time.perf_counter() - t0
    print(f'stage I on the GPU: {dt:.2f} s, {out["stagei_debug_details"]["b200"]}')
    _compare(out, ref, 1e-6)


def test_marker_layout_file_round_trip(cases, tmp_path):
    """The reference's call site passes no layout (mosh_head.py:242-244): Stage I reads cfg.dirs.marker_layout.fname."""
    from moshpp_b200 import synth
    from moshpp_b200.cfg import AttrDict
    case = cases('C2')
    fn = synth.write_marker_layout(str(tmp_path / 'layout.json'), case['marker_meta'])
    meta = product.load_marker_layout(fn, labels_map=None)
    ref = case['marker_meta']
    assert list(meta['marker_vids'].items()) == [(k, int(v)) for k, v in ref['marker_vids'].items()]
    assert dict(meta['marker_type']) == dict(ref['marker_type']) and meta['surface_model_type'] == ref['surface_model_type']
    for k in ref['marker_type_mask']:
        assert np.array_equal(meta['marker_type_mask'][k], ref['marker_type_mask'][k]) and meta['m2b_distance'][k] == ref['m2b_distance'][k]
