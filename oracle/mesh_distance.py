"""Point-to-triangle-mesh distance with derivatives (oracle; test infrastructure only) -- the surface term of Stage I,
the one native component of the reference.

Restates
  * scan2mesh/mesh_distance/sample2meshdist.h:67-207: ``pointPlane`` / ``pointLine`` / ``pointPoint`` with their
    derivatives wrt the sample and the three triangle vertices, dispatched by ``part`` (0 plane, 1..3 the edges ab / bc /
    ca, 4..6 the vertices a / b / c), under f = identity, square, Geman-McClure of the square (robust.h:14-52);
  * scan2mesh/mesh_distance/sample2meshdist.pyx:55-103 ``somedistance``: r[s] = f(dist(sample s, its nearest triangle));
  * the nearest-triangle query the reference gets from psbody.mesh's CGAL AABB tree (mesh_distance_main.py:346-376
    ``_AabbTree.nearest(v, nearest_part=True)``; external, absent here) as a brute-force closest-point-on-triangle search
    with the same (triangle, part) convention.

Pinned by oracle/_ref/libs2m.so: the reference header itself, compiled unmodified (oracle/build_ref.py;
tests/test_mesh_distance.py).
"""
from __future__ import annotations

import numpy as np

KIND_DISTANCE, KIND_SQUARED, KIND_GM = 0, 1, 2


def _f(kind, sigma, d):
    """(f(d), f'(d)) of robust.h: Identity, Square, Compose<GM, Square>."""
    if kind == KIND_DISTANCE:
        return d, 1.0
    if kind == KIND_SQUARED:
        return d * d, 2.0 * d
    s2 = sigma * sigma
    x2 = d * d
    val = s2 * x2 / (s2 + x2)
    dval = s2 * (1.0 / (s2 + x2)) - s2 * x2 / ((s2 + x2) ** 2)
    return val, dval * 2.0 * d


def point_plane(x, a, b, c, kind=KIND_DISTANCE, sigma=1.0):
    """Signed distance to the plane of abc = det(x-a, b-a, c-b) / |(b-a) x (c-b)| (sample2meshdist.h:67-100)."""
    A = np.stack([x - a, b - a, c - b])
    det = (A[0, 0] * (A[1, 1] * A[2, 2] - A[2, 1] * A[1, 2]) - A[1, 0] * (A[0, 1] * A[2, 2] - A[2, 1] * A[0, 2])
           + A[2, 0] * (A[0, 1] * A[1, 2] - A[1, 1] * A[0, 2]))
    ddetA = _adjugate(A)          # det(A) inv(A); column j = d det / d (row j of A)
    z, y = b - a, c - b
    s = np.linalg.norm(np.cross(z, y))
    ds_a = -(z * y.dot(y) - y * z.dot(y)) / s
    ds_c = (y * z.dot(z) - z * z.dot(y)) / s
    ds_b = -ds_a - ds_c
    fv, dfv = _f(kind, sigma, det / s)
    s2 = s * s
    dx = dfv * (ddetA[:, 0] / s)
    da = dfv * ((-ddetA[:, 0] - ddetA[:, 1]) / s - ds_a * (det / s2))
    db = dfv * ((ddetA[:, 1] - ddetA[:, 2]) / s - ds_b * (det / s2))
    dc = dfv * (ddetA[:, 2] / s - ds_c * (det / s2))
    return fv, dx, da, db, dc


def _adjugate(A):
    """det(A) inv(A) (sample2meshdist.h:50-65, det_times_inv3)."""
    return np.array([
        [A[2, 2] * A[1, 1] - A[2, 1] * A[1, 2], A[0, 2] * A[2, 1] - A[0, 1] * A[2, 2], A[0, 1] * A[1, 2] - A[0, 2] * A[1, 1]],
        [A[1, 2] * A[2, 0] - A[1, 0] * A[2, 2], A[0, 0] * A[2, 2] - A[0, 2] * A[2, 0], A[0, 2] * A[1, 0] - A[0, 0] * A[1, 2]],
        [A[1, 0] * A[2, 1] - A[1, 1] * A[2, 0], A[0, 1] * A[2, 0] - A[0, 0] * A[2, 1], A[0, 0] * A[1, 1] - A[0, 1] * A[1, 0]]])


def point_line(x, a, b, kind=KIND_DISTANCE, sigma=1.0):
    """|(x-a) x (x-b)| / |b-a| (sample2meshdist.h:134-158)."""
    w = np.cross(x - a, x - b)
    nw, nab = np.linalg.norm(w), np.linalg.norm(b - a)
    fv, dfv = _f(kind, sigma, nw / nab)
    r = w / (nw * nab)
    d = (b - a) * nw / nab ** 3
    return fv, dfv * np.cross(a - b, r), dfv * (np.cross(b - x, r) + d), dfv * (np.cross(x - a, r) - d)


def point_point(x, a, kind=KIND_DISTANCE, sigma=1.0):
    dist = np.linalg.norm(x - a)
    fv, dfv = _f(kind, sigma, dist)
    return fv, dfv * (x - a) / dist, dfv * (a - x) / dist


def tri(part, x, a, b, c, kind=KIND_DISTANCE, sigma=1.0):
    """``Distance<F>::tri`` (sample2meshdist.h:182-195): (value, dx, da, db, dc)."""
    z = np.zeros(3)
    if part == 0:
        return point_plane(x, a, b, c, kind, sigma)
    if part in (1, 2, 3):
        p, q = {1: (a, b), 2: (b, c), 3: (c, a)}[part]
        v, dx, dp, dq = point_line(x, p, q, kind, sigma)
        return {1: (v, dx, dp, dq, z), 2: (v, dx, z, dp, dq), 3: (v, dx, dq, z, dp)}[part]
    p = {4: a, 5: b, 6: c}[part]
    v, dx, dp = point_point(x, p, kind, sigma)
    return {4: (v, dx, dp, z, z), 5: (v, dx, z, dp, z), 6: (v, dx, z, z, dp)}[part]


def closest_on_triangles(x, A, B, C):
    """Squared distance from point x to each triangle (A[t], B[t], C[t]) and the part of the triangle the closest point
    lies on (0 interior, 1..3 edges ab / bc / ca, 4..6 vertices a / b / c).  Voronoi-region test, vectorised over t."""
    ab, ac, ap = B - A, C - A, x - A
    d1, d2 = (ab * ap).sum(1), (ac * ap).sum(1)
    bp = x - B
    d3, d4 = (ab * bp).sum(1), (ac * bp).sum(1)
    cp = x - C
    d5, d6 = (ab * cp).sum(1), (ac * cp).sum(1)
    vc = d1 * d4 - d3 * d2
    vb = d5 * d2 - d1 * d6
    va = d3 * d6 - d5 * d4
    T = len(A)
    part = np.zeros(T, dtype=np.int64)
    cl = np.zeros((T, 3))
    done = np.zeros(T, dtype=bool)

    def take(mask, p, pts):
        m = mask & ~done
        part[m] = p
        cl[m] = pts[m]
        done[m] = True

    with np.errstate(divide='ignore', invalid='ignore'):
        take((d1 <= 0) & (d2 <= 0), 4, A)
        take((d3 >= 0) & (d4 <= d3), 5, B)
        take((vc <= 0) & (d1 >= 0) & (d3 <= 0), 1, A + (d1 / (d1 - d3))[:, None] * ab)
        take((d6 >= 0) & (d5 <= d6), 6, C)
        take((vb <= 0) & (d2 >= 0) & (d6 <= 0), 3, A + (d2 / (d2 - d6))[:, None] * ac)
        take((va <= 0) & ((d4 - d3) >= 0) & ((d5 - d6) >= 0), 2, B + ((d4 - d3) / ((d4 - d3) + (d5 - d6)))[:, None] * (C - B))
        den = 1.0 / (va + vb + vc)
        take(np.ones(T, dtype=bool), 0, A + ab * (vb * den)[:, None] + ac * (vc * den)[:, None])
    return ((x - cl) ** 2).sum(1), part


def nearest(samples, verts, faces):
    """(nearest triangle, part) per sample -- what ``_AabbTree.nearest(v, nearest_part=True)`` returns."""
    A, B, C = verts[faces[:, 0]], verts[faces[:, 1]], verts[faces[:, 2]]
    tri_id = np.zeros(len(samples), dtype=np.int64)
    part = np.zeros(len(samples), dtype=np.int64)
    for s, x in enumerate(samples):
        d2, p = closest_on_triangles(x, A, B, C)
        t = int(np.argmin(d2))
        tri_id[s], part[s] = t, p[t]
    return tri_id, part


def somedistance(samples, verts, faces, kind=KIND_SQUARED, sigma=1.0, nearest_tri=None, nearest_part=None):
    """sample2meshdist.pyx:55-103: r [S], d r / d sample [S x 3], d r / d (a, b, c) [S x 9], and the (tri, part) used."""
    if nearest_tri is None:
        nearest_tri, nearest_part = nearest(samples, verts, faces)
    S = len(samples)
    r, dsample, dref = np.zeros(S), np.zeros((S, 3)), np.zeros((S, 9))
    for s in range(S):
        a, b, c = (verts[faces[nearest_tri[s], k]] for k in range(3))
        v, dx, da, db, dc = tri(int(nearest_part[s]), samples[s], a, b, c, kind, sigma)
        r[s], dsample[s], dref[s] = v, dx, np.concatenate([da, db, dc])
    return r, dsample, dref, nearest_tri, nearest_part
