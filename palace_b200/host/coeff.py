"""Coefficient context blobs with the exact layout the reference QFunctions read
(/root/reference/palace/fem/qfunctions/coeff/coeff_qf.h:7-43): an array of 8-byte
``union { int first; double second; }`` entries

    [n_attr][attr -> material (0-based) x n_attr][n_mat][material 3x3 coefficient, column-major x n_mat]

built the way ``PopulateCoefficientContext`` does (/root/reference/palace/fem/libceed/coefficient.cpp:51-130):
no coefficient -> one identity material and n_attr = 0; otherwise one extra all-zero material is
appended and unassigned (negative) attributes map to it; scalar properties are expanded to
diagonals; a pair context is the mass context followed by the curl context (coefficient.cpp:120-130).
"""
from __future__ import annotations

import numpy as np


def _set_int(blob, i, val):
    blob.view(np.int32)[2 * i] = val


def coeff_ctx(attr_mat=None, mat_coeff=None, a=1.0, transpose=False, dim=3) -> np.ndarray:
    cd = dim * dim
    if attr_mat is None:
        blob = np.zeros(2 + cd, dtype=np.float64)
        _set_int(blob, 0, 0)
        _set_int(blob, 1, 1)
        blob[2:2 + cd] = (a * np.eye(dim)).ravel(order="F")
        return blob
    attr_mat = np.asarray(attr_mat, dtype=np.int64)
    mat_coeff = np.asarray(mat_coeff, dtype=np.float64)  # [n_mat, dim, dim] or [n_mat] / [n_mat,1,1]
    n_mat = mat_coeff.shape[0]
    n_attr = attr_mat.size
    blob = np.zeros(2 + n_attr + cd * (n_mat + 1), dtype=np.float64)
    _set_int(blob, 0, n_attr)
    for i, k in enumerate(attr_mat):
        _set_int(blob, 1 + i, n_mat if k < 0 else int(k))
    _set_int(blob, 1 + n_attr, n_mat + 1)
    base = 2 + n_attr
    for k in range(n_mat):
        m = mat_coeff[k]
        if np.ndim(m) == 0 or m.size == 1:
            M = a * float(np.ravel(m)[0]) * np.eye(dim)
        else:
            M = a * (m.T if transpose else m)
        blob[base + cd * k: base + cd * (k + 1)] = M.ravel(order="F")
    return blob


def coeff_ctx_pair(first: np.ndarray, second: np.ndarray) -> np.ndarray:
    return np.concatenate([first, second])


def test_suite_coefficient(n_attr_global: int, kind: str = "matrix"):
    """The striped piecewise coefficient of the reference unit tests
    (/root/reference/test/unit/test-libceed.cpp:144-168): <= 4 materials, mat(k) = 0.1 off-diagonal and
    10 k + d + 1 on the diagonal; attribute i -> material i % n_mat."""
    n_mat = min(n_attr_global, 4)
    attr_mat = np.arange(n_attr_global) % n_mat
    if kind == "scalar":
        mc = np.array([[[10.0 * k + 1.0]] for k in range(n_mat)])
    else:
        mc = np.full((n_mat, 3, 3), 0.1)
        for k in range(n_mat):
            for d in range(3):
                mc[k, d, d] = 10.0 * k + (d + 1.0)
    return attr_mat, mc


def widen_scalar_ctx(blob1: np.ndarray) -> np.ndarray:
    """A coefficient context of dimension 1 (what PopulateCoefficientContext(1, Q) builds for the scalar curl of 2-D elements,
    integ/curlcurl.cpp:70) as the dimension-3 context of the same attribute -> material map with c * I materials."""
    ints = np.ascontiguousarray(blob1).view(np.int32)[::2]
    n_attr = int(ints[0])
    n_mat = int(ints[1 + n_attr])
    base = 2 + n_attr
    out = np.zeros(base + 9 * n_mat, dtype=np.float64)
    out[:base] = blob1[:base]
    for k in range(n_mat):
        out[base + 9 * k: base + 9 * (k + 1)] = (float(blob1[base + k]) * np.eye(3)).ravel(order="F")
    return out
