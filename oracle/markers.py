"""Marker attachment (oracle; test infrastructure only).  Reference: transformed_lm.py:45-162."""
from __future__ import annotations

import numpy as np
from sklearn.neighbors import NearestNeighbors

# transformed_lm.py:47-50 builds the candidate set from support_data/smplx_eyeballs.npz, whose
# content is exactly the vertex ids 9383..10474.
_EYEBALLS = set(range(9383, 10475))
NO_EYE_BALL_VIDS = sorted(set(np.arange(10474).tolist()).difference(_EYEBALLS))


def nrm(x):
    with np.errstate(invalid='ignore', divide='ignore'):
        return x / np.sqrt(np.sum(x ** 2, axis=1)).reshape((-1, 1))


class TransformedCoeffs:
    """transformed_lm.py:59-113: 8-NN (kd-tree) local frames on the canonical body."""

    def __init__(self, can_body: np.ndarray, markers_latent: np.ndarray):
        can_body = np.asarray(can_body, dtype=np.float64)
        markers_latent = np.asarray(markers_latent, dtype=np.float64)
        if len(can_body) == 10475:
            cand = can_body[NO_EYE_BALL_VIDS]
        else:
            cand = can_body
        nn = NearestNeighbors(algorithm='kd_tree', n_neighbors=8).fit(cand)
        _, closest = nn.kneighbors(markers_latent)
        self.closest = np.vstack(closest)
        diff = (markers_latent - can_body[self.closest[:, 0]]).reshape((-1, 3))
        e1 = can_body[self.closest[:, 1]] - can_body[self.closest[:, 0]]
        e2 = can_body[self.closest[:, 2]] - can_body[self.closest[:, 0]]
        f1 = nrm(e1)
        counter = 3
        while np.isnan(nrm(np.cross(e1, e2)).sum()) and counter < self.closest.shape[0]:
            e2 = can_body[self.closest[:, counter]] - can_body[self.closest[:, 0]]
            counter += 1
        self.closest[:, 2] = self.closest[:, counter - 1]
        f2 = nrm(np.cross(e1, e2))
        f3 = np.cross(f1, f2)
        self.coefs = np.hstack([(diff * f1).sum(1, keepdims=True), (diff * f2).sum(1, keepdims=True),
                                (diff * f3).sum(1, keepdims=True)])

    @property
    def vids(self):
        """Unique vertex ids the markers depend on (rows the lean oracle evaluates)."""
        return np.unique(self.closest[:, :3].reshape(-1))


def _N(u):
    """d nrm(u) / du = (I - uh uh^T) / |u|."""
    n = np.linalg.norm(u)
    uh = u / n
    return (np.eye(3) - np.outer(uh, uh)) / n


def _skew(w):
    return np.array([[0.0, -w[2], w[1]], [w[2], 0.0, -w[0]], [-w[1], w[0], 0.0]])


def transformed_lms(tc: TransformedCoeffs, v0, v1, v2, want_jac=False):
    """transformed_lm.py:130-159 on the posed triples (M x 3 each).  The collinear re-selection of the
    reference (lines 143-147) never triggers for non-degenerate posed triangles and is not restated.
    Returns markers M x 3 and, if want_jac, local Jacobians M x 3 x 9 wrt (v0, v1, v2)."""
    e1 = v1 - v0
    e2 = v2 - v0
    f1 = nrm(e1)
    f2 = nrm(np.cross(e1, e2))
    f3 = np.cross(f1, f2)
    k = tc.coefs
    out = v0 + k[:, 0:1] * f1 + k[:, 1:2] * f2 + k[:, 2:3] * f3
    if not want_jac:
        return out
    M = v0.shape[0]
    loc = np.zeros((M, 3, 9))
    for i in range(M):
        n = np.cross(e1[i], e2[i])
        df1_de1 = _N(e1[i])
        df2_de1 = _N(n).dot(-_skew(e2[i]))
        df2_de2 = _N(n).dot(_skew(e1[i]))
        df3_de1 = -_skew(f2[i]).dot(df1_de1) + _skew(f1[i]).dot(df2_de1)
        df3_de2 = _skew(f1[i]).dot(df2_de2)
        d_e1 = k[i, 0] * df1_de1 + k[i, 1] * df2_de1 + k[i, 2] * df3_de1
        d_e2 = k[i, 1] * df2_de2 + k[i, 2] * df3_de2
        loc[i, :, 0:3] = np.eye(3) - d_e1 - d_e2
        loc[i, :, 3:6] = d_e1
        loc[i, :, 6:9] = d_e2
    return out, loc
