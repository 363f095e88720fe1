"""Rotation maps and the Procrustes initialisation (oracle; test infrastructure only).

Reference: rigid_transformations.py:39-83 (Arun/Procrustes + cv2.Rodrigues), and the axis-angle
convention of cv2.Rodrigues used by psbody.smpl's ``lrotmin`` pose features.
"""
from __future__ import annotations

import numpy as np

_E = np.zeros((3, 3, 3))
_E[0, 1, 2], _E[0, 2, 1] = -1.0, 1.0
_E[1, 0, 2], _E[1, 2, 0] = 1.0, -1.0
_E[2, 0, 1], _E[2, 1, 0] = -1.0, 1.0     # _E[k] = [e_k]_x


def skew(w):
    return np.array([[0.0, -w[2], w[1]], [w[2], 0.0, -w[0]], [-w[1], w[0], 0.0]])


def rodrigues_coeffs(t2: float):
    """a = sin t / t, b = (1-cos t)/t^2 and c1 = a'(t)/t, c2 = b'(t)/t as functions of t^2."""
    if t2 < 1e-4:
        a = 1.0 - t2 / 6.0 + t2 * t2 / 120.0
        b = 0.5 - t2 / 24.0 + t2 * t2 / 720.0
        c1 = -1.0 / 3.0 + t2 / 30.0 - t2 * t2 / 840.0
        c2 = -1.0 / 12.0 + t2 / 180.0 - t2 * t2 / 6720.0
    else:
        t = np.sqrt(t2)
        s, c = np.sin(t), np.cos(t)
        a = s / t
        b = (1.0 - c) / t2
        c1 = (t * c - s) / (t2 * t)
        c2 = (t * s - 2.0 * (1.0 - c)) / (t2 * t2)
    return a, b, c1, c2


def rodrigues(rv, want_jac: bool = False):
    """exp([rv]_x) with the cv2.Rodrigues convention; optionally dR[k] = dR/d rv_k (3 x 3 x 3)."""
    rv = np.asarray(rv, dtype=np.float64)
    K = skew(rv)
    K2 = K @ K
    a, b, c1, c2 = rodrigues_coeffs(float(rv @ rv))
    R = np.eye(3) + a * K + b * K2
    if not want_jac:
        return R
    dR = np.zeros((3, 3, 3))
    for k in range(3):
        dR[k] = c1 * rv[k] * K + a * _E[k] + c2 * rv[k] * K2 + b * (_E[k] @ K + K @ _E[k])
    return R, dR


def rodrigues_inv(R):
    """log map R -> axis-angle, angle in [0, pi] (what cv2.Rodrigues returns for a 3x3 input;
    rigid_transformations.py:82)."""
    R = np.asarray(R, dtype=np.float64)
    u, _, vt = np.linalg.svd(R)
    R = u @ vt
    r = np.array([R[2, 1] - R[1, 2], R[0, 2] - R[2, 0], R[1, 0] - R[0, 1]])
    s = np.sqrt((r @ r) * 0.25)
    c = np.clip((np.trace(R) - 1.0) * 0.5, -1.0, 1.0)
    theta = np.arccos(c)
    if s < 1e-5:
        if c > 0:
            return np.zeros(3)
        t = (R[0, 0] + 1) * 0.5
        rx = np.sqrt(max(t, 0.0))
        t = (R[1, 1] + 1) * 0.5
        ry = np.sqrt(max(t, 0.0)) * (-1.0 if R[0, 1] < 0 else 1.0)
        t = (R[2, 2] + 1) * 0.5
        rz = np.sqrt(max(t, 0.0)) * (-1.0 if R[0, 2] < 0 else 1.0)
        if abs(rx) < abs(ry) and abs(rx) < abs(rz) and (R[1, 2] > 0) != (ry * rz > 0):
            rz = -rz
        v = np.array([rx, ry, rz])
        return v * (theta / np.linalg.norm(v))
    return r * (theta / (2.0 * s))


def rigid_landmark_transform(a, b):
    """(R, T) with R a + T ~= b for 3 x N arrays (Arun et al. 1987); rigid_transformations.py:39-69."""
    assert a.shape[0] == 3 and b.shape[0] == 3
    b = np.where(np.isnan(b), a, b)
    a_mean = np.mean(a, axis=1).reshape((-1, 1))
    b_mean = np.mean(b, axis=1).reshape((-1, 1))
    c = (a - a_mean).dot((b - b_mean).T)
    u, _, v = np.linalg.svd(c, full_matrices=False)
    v = v.T
    R = v.dot(u.T)
    if np.linalg.det(R) < 0:
        v[:, 2] = -v[:, 2]
        R = v.dot(u.T)
    T = (b_mean - R.dot(a_mean)).reshape((-1, 1))
    return R, T


def perform_rigid_adjustment(markers_sim, markers_obs):
    """Returns (root axis-angle, trans) as assigned at rigid_transformations.py:72-83."""
    R, T = rigid_landmark_transform(np.asarray(markers_sim).T, np.asarray(markers_obs).T)
    return rodrigues_inv(R), T.ravel()
