"""Distance of points to a triangle mesh, with derivatives, on the GPU -- the surface term of MoSh++ Stage I.

Replaces the reference's only native code: ``sample2meshdist.pyx:55-103`` (``somedistance`` over
``sample2meshdist.h:67-207``) plus the nearest (triangle, part) query of psbody.mesh's AABB tree
(``mesh_distance_main.py:346-376``).  ``mesh_distance`` returns what ``somedistance`` returns -- the residual and its
derivatives wrt the samples and wrt the reference vertices (``as_sparse`` builds the same scipy matrices) -- from
``libmosh2.so`` (C-ABI ``mosh2_mesh_distance``); there is no CPU path.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional

import numpy as np

from . import lib as _lib

KINDS = {'distance': 0, 'squared': 1, 'gm': 2}


def mesh_distance(sample_verts: np.ndarray, reference_verts: np.ndarray, reference_faces: np.ndarray, *, kind=1, sigma: float = 1.0,
                  nearest_tri: Optional[np.ndarray] = None, nearest_part: Optional[np.ndarray] = None, device: int = 0) -> Dict:
    """r[s] = f(dist(sample s, mesh)); f by ``kind`` ('distance' | 'squared' | 'gm' or 0 | 1 | 2; ``sigma`` for 'gm').
    Returns value [S], tri [S], part [S] (0 plane, 1..3 edges ab/bc/ca, 4..6 vertices a/b/c), d_sample [S,3], d_tri [S,9]
    and the device time ``kernel_ms``."""
    L = _lib.load_library()
    k = KINDS[kind] if isinstance(kind, str) else int(kind)
    s = np.ascontiguousarray(sample_verts, dtype=np.float64).reshape(-1, 3)
    v = np.ascontiguousarray(reference_verts, dtype=np.float64).reshape(-1, 3)
    f = np.ascontiguousarray(reference_faces, dtype=np.int32).reshape(-1, 3)
    S = len(s)
    out = dict(value=np.zeros(S), tri=np.zeros(S, dtype=np.int32), part=np.zeros(S, dtype=np.int32),
               d_sample=np.zeros((S, 3)), d_tri=np.zeros((S, 9)))
    o = _lib.MeshDistanceOut(_lib._ptr(out['value'], _lib._f64p), _lib._ptr(out['tri'], _lib._i32p), _lib._ptr(out['part'], _lib._i32p),
                             _lib._ptr(out['d_sample'], _lib._f64p), _lib._ptr(out['d_tri'], _lib._f64p))
    nt = nprt = None
    if nearest_tri is not None:
        nt = np.ascontiguousarray(nearest_tri, dtype=np.int32)
        nprt = np.ascontiguousarray(nearest_part, dtype=np.int32)
    ms = C.c_float()
    rc = L.mosh2_mesh_distance(device, k, float(sigma), S, _lib._ptr(s, _lib._f64p), len(v), _lib._ptr(v, _lib._f64p), len(f),
                               _lib._ptr(f, _lib._i32p), _lib._ptr(nt, _lib._i32p) if nt is not None else None,
                               _lib._ptr(nprt, _lib._i32p) if nprt is not None else None, C.byref(o), C.byref(ms))
    if rc != 0:
        raise _lib.Mosh2Error(f'mosh2_mesh_distance failed ({rc}): {L.mosh2_last_error().decode()}')
    out['kernel_ms'] = float(ms.value)
    return out


def as_sparse(out: Dict, reference_faces: np.ndarray, n_reference_verts: int):
    """(Dr_refv [S x 3V], Dr_samplev [S x 3S]) as scipy CSC matrices -- the layout ``somedistance`` returns
    (sample2meshdist.pyx:84-101)."""
    import scipy.sparse as sp
    f = np.asarray(reference_faces).reshape(-1, 3)
    S = len(out['value'])
    rows = np.repeat(np.arange(S), 9)
    cols = (3 * f[out['tri']][:, :, None] + np.arange(3)[None, None, :]).reshape(-1)
    Dr_ref = sp.coo_matrix((out['d_tri'].reshape(-1), (rows, cols)), shape=(S, 3 * n_reference_verts)).tocsc()
    js = np.arange(3 * S)
    Dr_sample = sp.coo_matrix((out['d_sample'].reshape(-1), (js // 3, js)), shape=(S, 3 * S)).tocsc()
    return Dr_ref, Dr_sample
