"""Pins against the UNMODIFIED reference: vectors produced by the reference's own importable modules
(tests/golden/make_reference_vectors.py, run where /root/reference exists) -- `moshpp.rigid_transformations` for the
first-frame rigid adjustment and `moshpp.tools.c3d` for the c3d metadata (row f-1).  The rest
of the Stage-II path cannot be imported (chumpy / psbody / ezc3d absent) and stays pinned by the oracle's own checks."""
import os

import numpy as np
import pytest

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


@pytest.mark.parametrize('proc', ['dec', 'mips'])
def test_c3d_dec_and_mips_files_as_the_reference_parser_reads_them(tmp_path, proc):
    """The same content in the DEC (VAX F-floating numbers) and SGI/MIPS (big-endian) processor formats: the file our writer
    produces is the one the reference's parser (tools/c3d.py:35-100,368-424, unmodified) read rate, scale, counts and labels
    from, and our reader returns the data."""
    g = np.load(os.path.join(GOLD, 'ref_c3d.npz'))
    fn = str(tmp_path / f'{proc}.c3d')
    labels = [str(s) for s in g['in_labels']]
    c3d_io.write_c3d(fn, g['data'], labels, frame_rate=120.0, processor=proc)
    with open(fn, 'rb') as h:
        assert np.array_equal(np.frombuffer(h.read(), dtype=np.uint8), g[f'{proc}_file_bytes'])
    assert float(g[f'{proc}_point_rate']) == 120.0 and float(g[f'{proc}_point_scale']) == -1.0
    assert int(g[f'{proc}_point_used']) == len(labels)
    assert int(g[f'{proc}_last_frame']) - int(g[f'{proc}_first_frame']) + 1 == g['data'].shape[0]
    assert [str(s) for s in g[f'{proc}_labels']] == labels
    pts, lab, rate = c3d_io.read_c3d(fn)
    assert lab == labels and rate == 120.0
    assert np.allclose(pts, g['data'], rtol=0, atol=1e-3, equal_nan=True)


def test_dec_float_conversion_equals_the_reference():
    """VAX F-floating -> IEEE: the reference's own converters (tools/c3d.py:115-190, array and scalar form) on 455 numbers."""
    g = np.load(os.path.join(GOLD, 'ref_c3d.npz'))
    mine = c3d_io.dec_to_ieee(g['dec_bytes'].tobytes())
    ref = g['dec_as_ieee_by_reference']
    normal = np.abs(ref) > 1e-30                       # (the reference's exponent decrement is undefined on zero)
    assert normal.sum() >= 450 and np.array_equal(mine[normal], ref[normal])
    assert np.array_equal(mine[:64][normal[:64]], g['dec_scalar_by_reference'][normal[:64]])
    assert np.array_equal(c3d_io.dec_to_ieee(c3d_io.ieee_to_dec(ref[normal])), ref[normal])


# --------------------------------------------------------------------------------------------------------------------
# rows a5, a6, a7: vectors emitted by the UNMODIFIED prior/gmm_prior_ch.py and transformed_lm.py running over the
# forward-only chumpy stand-in tests/golden/ref_shim (make_reference_vectors.py: prior_and_marker_vectors)
# --------------------------------------------------------------------------------------------------------------------
def _prior_file(tmp_path, g):
    import pickle
    fn = str(tmp_path / 'prior.pkl')
    with open(fn, 'wb') as f:
        pickle.dump(dict(covars=g['covars'], means=g['means'], weights=g['weights']), f)
    return fn


def test_oracle_prior_equals_reference(tmp_path):
    """create_gmm_body_prior (gmm_prior_ch.py:107-134) and MaxMixtureComplete (42-72): normalised weights, Cholesky
    factors, arg-min component and the D+1 residual on probe poses that select every component."""
    from oracle import prior as oprior
    g = np.load(os.path.join(GOLD, 'ref_prior.npz'))
    fn = _prior_file(tmp_path, g)
    for tag, excl in (('63', True), ('69', False)):
        mm = oprior.create_gmm_body_prior(fn, exclude_hands=excl)
        assert np.abs(mm.precs - g[f'chols_{tag}']).max() < 1e-10 * np.abs(g[f'chols_{tag}']).max()
        assert np.allclose(mm.weights, g[f'weights_{tag}'].ravel(), rtol=1e-12, atol=0)
        assert np.array_equal(mm.means, g[f'means_{tag}'])
        for x, r_ref, k_ref in zip(g[f'x_{tag}'], g[f'r_{tag}'], g[f'k_{tag}']):
            k, _ = mm.select(x)
            assert k == int(k_ref)
            assert np.abs(mm.r(x) - r_ref).max() < 1e-9 * max(1.0, np.abs(r_ref).max())
        assert len(set(g[f'k_{tag}'].tolist())) == len(mm.weights)          # every component was exercised


def test_device_prior_constants_equal_reference(tmp_path):
    """What the device consumes (pack.create_gmm_body_prior: Q = .5 inv(cov), -log w) reproduces the reference's residual:
    sum r[:-1]^2 = (x - mu_k)^T Q_k (x - mu_k), r[-1]^2 = -log w_k, and the same arg-min component."""
    from moshpp_b200 import pack
    g = np.load(os.path.join(GOLD, 'ref_prior.npz'))
    fn = _prior_file(tmp_path, g)
    for tag, excl in (('63', True), ('69', False)):
        bp = pack.create_gmm_body_prior(fn, exclude_hands=excl)
        L = g[f'chols_{tag}']
        assert np.abs(bp.Q - 0.5 * L @ np.transpose(L, (0, 2, 1))).max() < 1e-9 * np.abs(bp.Q).max()
        for x, r_ref, k_ref in zip(g[f'x_{tag}'], g[f'r_{tag}'], g[f'k_{tag}']):
            dx = x[None] - bp.means
            q = np.einsum('ki,kij,kj->k', dx, bp.Q, dx) + bp.neglogw            # what the kernel minimises over k
            k = int(np.argmin(q))
            assert k == int(k_ref)
            assert abs(q[k] - (r_ref ** 2).sum()) < 1e-9 * max(1.0, (r_ref ** 2).sum())
            assert abs(bp.neglogw[k] - r_ref[-1] ** 2) < 1e-10 * max(1.0, r_ref[-1] ** 2)


def test_marker_attachment_equals_reference():
    """TransformedCoeffs (transformed_lm.py:59-113) incl. the SMPL-X eyeball exclusion and the collinear-neighbour
    fallback, TransformedLms (130-159): the oracle's restatement and the product's host pack against the reference."""
    from moshpp_b200 import pack
    from oracle import markers as omk
    g = np.load(os.path.join(GOLD, 'ref_lms.npz'))
    for tag in ('smplh', 'smplx', 'line'):
        can, mk = g[f'can_{tag}'], g[f'markers_latent_{tag}']
        ref_closest, ref_coefs = g[f'closest_{tag}'][:, :3], g[f'coefs_{tag}']
        tc = omk.TransformedCoeffs(can, mk)
        assert np.array_equal(tc.closest[:, :3], ref_closest)
        assert np.abs(tc.coefs - ref_coefs).max() < 1e-12
        closest, coefs = pack.attach_markers(can, mk)                           # brute-force 8-NN of the product
        assert np.array_equal(closest, ref_closest)
        assert np.abs(coefs - ref_coefs).max() < 1e-12
        posed = g[f'posed_{tag}']
        sim = omk.transformed_lms(tc, posed[ref_closest[:, 0]], posed[ref_closest[:, 1]], posed[ref_closest[:, 2]])
        assert np.abs(sim - g[f'markers_{tag}']).max() < 1e-12
    assert (g['closest_smplx'] < 9383).all()                                    # eyeball vertices were never chosen
    assert g['closest_line'][0, 2] != 2                                         # the collinear third neighbour was swapped
