"""Pins of the float64 oracle: the reference ships no tests, so the oracle is checked
against independent implementations -- torch.autograd, cv2.Rodrigues, scipy.optimize."""
import numpy as np
import pytest
import torch

from oracle import dogleg, lbs, markers, prior, rigid, stageii


def test_rodrigues_matches_cv2():
    import cv2
    rng = np.random.default_rng(0)
    for s in [1e-7, 1e-3, 0.05, 0.3, 1.0, 2.0, 3.1]:
        rv = rng.standard_normal(3)
        rv *= s / np.linalg.norm(rv)
        R, dR = rigid.rodrigues(rv, True)
        Rc, Jc = cv2.Rodrigues(rv)
        assert np.abs(R - Rc).max() < 1e-12
        assert np.abs(dR.reshape(3, 9) - Jc).max() < 1e-8
        if s > 1e-5:
            assert np.abs(rigid.rodrigues_inv(R) - rv).max() < 1e-9
            assert np.abs(rigid.rodrigues_inv(R) - cv2.Rodrigues(R)[0].ravel()).max() < 1e-9


def test_rigid_landmark_transform_recovers_motion():
    rng = np.random.default_rng(1)
    a = rng.standard_normal((3, 30))
    R = rigid.rodrigues(np.array([0.3, -1.1, 0.5]))
    T = np.array([[0.4], [-0.2], [1.5]])
    Rh, Th = rigid.rigid_landmark_transform(a, R @ a + T)
    assert np.abs(Rh - R).max() < 1e-12 and np.abs(Th - T).max() < 1e-12
    # reflection case keeps a proper rotation
    b = np.diag([1, 1, -1.0]) @ a
    Rh, _ = rigid.rigid_landmark_transform(a, b)
    assert np.linalg.det(Rh) > 0


def _torch_forward(m, rows, pose, betas, trans):
    """Independent float64 torch restatement of the SMPL forward (used only for autograd)."""
    def rod(w):
        th = torch.sqrt((w * w).sum() + 1e-300)
        k = w / th
        K = torch.zeros(3, 3, dtype=torch.float64)
        K[0, 1], K[0, 2], K[1, 0], K[1, 2], K[2, 0], K[2, 1] = -k[2], k[1], k[2], -k[0], -k[1], k[0]
        return torch.eye(3, dtype=torch.float64) + torch.sin(th) * K + (1 - torch.cos(th)) * (K @ K)
    T = lambda a: torch.tensor(a, dtype=torch.float64)
    nj = m.n_joints
    if m.selected_components.shape[0]:
        full = torch.cat([pose[:m.body_dof], T(m.hands_mean) + pose[m.body_dof:] @ T(m.selected_components)])
    else:
        full = pose
    vs_all = T(m.v_template) + torch.einsum('vcb,b->vc', T(m.shapedirs[:, :, :len(betas)]), betas)
    J = T(m.J_regressor) @ vs_all
    R = [rod(full[3 * j:3 * j + 3]) for j in range(nj)]
    pf = torch.cat([(R[j] - torch.eye(3, dtype=torch.float64)).reshape(-1) for j in range(1, nj)])
    vp = vs_all[rows] + torch.einsum('vcp,p->vc', T(m.posedirs[rows]), pf)
    Rg, tg = [R[0]], [J[0]]
    for j in range(1, nj):
        a = int(m.parents[j])
        Rg.append(Rg[a] @ R[j])
        tg.append(tg[a] + Rg[a] @ (J[j] - J[a]))
    W = T(m.weights[rows])
    out = torch.zeros(len(rows), 3, dtype=torch.float64)
    for j in range(nj):
        out = out + W[:, j:j + 1] * ((vp - J[j]) @ Rg[j].T + tg[j])
    return out + trans


@pytest.mark.parametrize('name', ['C1', 'C2', 'C3', 'C4'])
def test_lbs_jacobian_matches_autograd(cases, name):
    case = cases(name)
    sol = stageii.StageIISolver(case['cfg'], case['markers_latent'], case['latent_labels'], case['betas'],
                                case['marker_meta'], mode='lean')
    m = sol.model
    rng = np.random.default_rng(3)
    pose = 0.3 * rng.standard_normal(m.pose_size)
    betas = sol.betas.copy()
    if sol.nd:
        betas[sol.dmpl_ids] = 0.5 * rng.standard_normal(sol.nd)
    trans = rng.standard_normal(3)
    rows = sol.vids[:40]
    sub = lbs.LBS(m, rows)
    v, dv_pose, dv_beta = sub(pose, betas, trans, True, beta_ids=sol.dmpl_ids)
    tp = torch.tensor(pose, dtype=torch.float64)
    tb = torch.tensor(betas, dtype=torch.float64)
    tt = torch.tensor(trans, dtype=torch.float64)
    vt = _torch_forward(m, rows, tp, tb, tt)
    assert np.abs(vt.numpy() - v).max() < 1e-12
    Jp = torch.autograd.functional.jacobian(lambda p: _torch_forward(m, rows, p, tb, tt), tp).numpy()
    assert np.abs(Jp - dv_pose).max() < 1e-9 * max(1.0, np.abs(Jp).max())
    if sol.nd:
        Jb = torch.autograd.functional.jacobian(lambda b: _torch_forward(m, rows, tp, b, tt), tb).numpy()
        assert np.abs(Jb[:, :, sol.dmpl_ids] - dv_beta).max() < 1e-9


def test_marker_local_jacobian_finite_difference():
    rng = np.random.default_rng(5)
    can = rng.standard_normal((200, 3))
    lat = can[rng.choice(200, 9, replace=False)] + 0.01 * rng.standard_normal((9, 3))
    tc = markers.TransformedCoeffs(can, lat)
    c = tc.closest
    v = [can[c[:, k]] + 0.05 * rng.standard_normal((9, 3)) for k in range(3)]
    mk, loc = markers.transformed_lms(tc, v[0], v[1], v[2], True)
    # on the canonical body itself the attachment reproduces the latent markers
    assert np.abs(markers.transformed_lms(tc, can[c[:, 0]], can[c[:, 1]], can[c[:, 2]]) - lat).max() < 1e-12
    eps = 1e-6
    for k in range(3):
        for ax in range(3):
            vp = [x.copy() for x in v]
            vm = [x.copy() for x in v]
            vp[k][:, ax] += eps
            vm[k][:, ax] -= eps
            fd = (markers.transformed_lms(tc, *vp) - markers.transformed_lms(tc, *vm)) / (2 * eps)
            assert np.abs(fd - loc[:, :, 3 * k + ax]).max() < 1e-7


def test_full_mesh_mode_selects_the_same_rows(cases):
    case = cases('C2')
    a = stageii.StageIISolver(case['cfg'], case['markers_latent'], case['latent_labels'], case['betas'], case['marker_meta'], mode='lean')
    b = stageii.StageIISolver(case['cfg'], case['markers_latent'], case['latent_labels'], case['betas'], case['marker_meta'], mode='reference_cost')
    a.pose[:] = case['gt_pose'][2]; b.pose[:] = case['gt_pose'][2]
    a.trans[:] = case['gt_trans'][2]; b.trans[:] = case['gt_trans'][2]
    ea, eb = a.evaluate(True), b.evaluate(True)
    assert np.abs(ea['markers'] - eb['markers']).max() < 1e-13
    assert np.abs(ea['dm_pose'] - eb['dm_pose']).max() < 1e-12


def test_max_mixture_prior_forms_agree(cases):
    """The product uses Q_k = .5 inv(cov_k); the reference's residual is sqrt(.5)(x-mu) chol(inv(cov_k))."""
    from moshpp_b200 import pack
    case = cases('C2')
    fn = case['cfg'].moshpp.pose_body_prior_fname
    mm = prior.create_gmm_body_prior(fn, exclude_hands=True)
    bp = pack.create_gmm_body_prior(fn, exclude_hands=True)
    rng = np.random.default_rng(2)
    for _ in range(5):
        x = 0.3 * rng.standard_normal(63)
        r = mm.r(x)
        q = np.array([(x - mu) @ Q @ (x - mu) + nl for mu, Q, nl in zip(bp.means, bp.Q, bp.neglogw)])
        k, _ = mm.select(x)
        assert int(np.argmin(q)) == k
        assert abs(q.min() - (r ** 2).sum()) < 1e-9 * max(1.0, q.min())
        J = mm.dr_wrt_x(x)
        assert np.abs(J.T @ J - bp.Q[k]).max() < 1e-9 * np.abs(bp.Q[k]).max()
        assert np.abs(J.T @ r - bp.Q[k] @ (x - bp.means[k])).max() < 1e-8 * np.abs(J.T @ r).max()


def test_dogleg_reaches_scipy_optimum():
    from scipy.optimize import least_squares
    rng = np.random.default_rng(7)
    A = rng.standard_normal((40, 6))
    b = rng.standard_normal(40)

    def f(x, want_jac):
        r = np.tanh(A @ x) - b + 0.1 * np.concatenate([x, np.zeros(34)])
        if not want_jac:
            return r
        J = (1 - np.tanh(A @ x) ** 2)[:, None] * A
        J[:6] += 0.1 * np.eye(6)
        return r, J
    x, st = dogleg.minimize_dogleg(f, np.zeros(6), e_3=0.0, delta_0=0.5, maxiter=200)
    ref = least_squares(lambda x: f(x, False), np.zeros(6), jac=lambda x: f(x, True)[1], xtol=1e-14, ftol=1e-14, gtol=1e-14)
    assert np.abs(x - ref.x).max() < 1e-6
    # e_3 stops as soon as a step improves the SSE by less than the ratio
    x3, st3 = dogleg.minimize_dogleg(f, np.zeros(6), e_3=1e-2, delta_0=0.5, maxiter=200)
    assert st3.stop_reason == 'small improvement' and st3.iterations <= st.iterations


def test_oracle_solution_is_a_local_optimum_of_step2(cases):
    """scipy's trust-region solver started at the oracle's Step-2 solution does not move it (much)."""
    from scipy.optimize import least_squares
    case = cases('C4')
    sol = stageii.StageIISolver(case['cfg'], case['markers_latent'], case['latent_labels'], case['betas'], case['marker_meta'])
    from conftest import dense_obs
    obs, vis = dense_obs(case)
    frames = [(np.nonzero(vis[t])[0], obs[t][vis[t]]) if vis[t].any() else None for t in range(3)]
    out = sol.solve_range(frames)
    vi, ob = frames[2]
    # rebuild the Step-2 objective of the last frame at the solution
    wt = sol.wts
    anneal = 1.0 + (sol.n_markers - len(vi)) / sol.n_markers * wt['stageii_wt_annealing']
    terms = [['data', wt['stageii_wt_data'] * 46 / len(vi)], ['velo', (wt['stageii_wt_velo'], 2 * out[1]['pose'] - out[0]['pose'])],
             ['poseH', wt['stageii_wt_poseH'] * anneal]]
    obj = stageii._Objective(sol, ob, vi, terms, sol.step2_ids, False)
    x0 = obj.x0()
    r0 = obj(x0, False)
    ref = least_squares(lambda x: obj(x, False), x0, jac=lambda x: obj(x, True)[1], xtol=1e-12, ftol=1e-12)
    assert (r0 ** 2).sum() <= (ref.fun ** 2).sum() * 1.02      # the e_3 = 1 % stop rule leaves at most ~1 %
