"""Boundary integrators on TETRAHEDRAL meshes (surface impedance / lumped-port / absorbing terms of BASELINE config 3's kind):
caller-side stand-in for `a.AddBoundaryIntegrator<VectorFEMassIntegrator>(fb)` on triangles
(/root/reference/palace/models/spaceoperator.cpp:300-303; the reference assembles them on ND triangle elements with the dim = 2,
space_dim = 3 QFunctions, fem/qfunctions/32/hcurl_32_qf.h).

No triangle element is needed to feed the dense-basis operator: the tangential trace of the volume space on a boundary face IS the
face space, so a boundary face enters as an "element" whose tables are the PARENT tetrahedron's tables at the face's quadrature
points, whose geometry factor is the parent's 3-D J^-T there, whose weight is the surface measure w |J e_s x J e_t|, and whose
coefficient is c (I - n n^T) -- the 3-D mass map w A^T C A u then integrates c u_t . v_t over the face; basis functions of other
faces / edges / the interior have no tangential trace there and drop out by themselves. One dense operator per LOCAL face index
(its points differ), all on the parent's restriction rows. Planar faces (straight-sided tets): the normal is constant per face and
rides in the per-face material."""
from __future__ import annotations

import dataclasses

import numpy as np

from . import tetspace as ts

REF_VERTS = np.array([[0.0, 0.0, 0.0], [1.0, 0.0, 0.0], [0.0, 1.0, 0.0], [0.0, 0.0, 1.0]])


def tri_quadrature(degree: int):
    """Conical Gauss-Jacobi rule on the reference triangle (s, t >= 0, s + t <= 1), exact to `degree`; weights sum to 1/2."""
    from scipy.special import roots_jacobi

    n = degree // 2 + 1
    x1, w1 = roots_jacobi(n, 1.0, 0.0)
    x2, w2 = roots_jacobi(n, 0.0, 0.0)
    a, wa = 0.5 * (x1 + 1), w1 / 4.0
    b, wb = 0.5 * (x2 + 1), w2 / 2.0
    A, B = np.meshgrid(a, b, indexing="ij")
    W = wa[:, None] * wb[None, :]
    return np.stack([A.ravel(), (B * (1 - A)).ravel()], axis=1), W.ravel()


@dataclasses.dataclass
class TetBdrGroup:
    local_face: int          # the parent's vertex the face is opposite to
    elems: np.ndarray        # parent tetrahedra
    interp: np.ndarray       # [3][Q2][P] parent tables at the face points
    qdata: np.ndarray        # [nf][11][Q2] {face attribute, surface weight, parent J^-T}
    idx: np.ndarray          # [nf][P] parent restriction rows
    curl_orient: np.ndarray  # [nf][P][3]
    normals: np.ndarray      # [nf][3] unit normals
    areas: np.ndarray        # [nf]


def boundary_groups(mesh: ts.TetMesh, nd: ts.TetSpace, select=None, degree=None):
    """The boundary faces of a straight-sided tet mesh (faces with one adjacent tet; `select(centroid, normal)` filters them),
    grouped by local face index; face attributes number the selected faces 1 ... N over all groups (one material per face)."""
    el = ts.nd_tet_element(nd.p)
    pts2, w2 = tri_quadrature(degree if degree is not None else 2 * nd.p)
    count = {}
    for e in range(mesh.ne):
        v = mesh.elems[e]
        for f in range(4):
            key = tuple(sorted(int(v[t]) for t in range(4) if t != f))
            count.setdefault(key, []).append((e, f))
    groups, next_attr = [], 1
    for f in range(4):
        others = [t for t in range(4) if t != f]
        es, et = REF_VERTS[others[1]] - REF_VERTS[others[0]], REF_VERTS[others[2]] - REF_VERTS[others[0]]
        rpts = REF_VERTS[others[0]][None] + pts2[:, :1] * es[None] + pts2[:, 1:] * et[None]
        interp, _ = el.tabulate(rpts)
        sel, normals, areas = [], [], []
        for key, lst in count.items():
            if len(lst) != 1 or lst[0][1] != f:
                continue
            e = lst[0][0]
            X = mesh.verts[mesh.elems[e]]
            J = np.stack([X[1] - X[0], X[2] - X[0], X[3] - X[0]], axis=1)
            nvec = np.cross(J @ es, J @ et)
            n = nvec / np.linalg.norm(nvec)
            if select is not None and not select(X[others].mean(axis=0), n):
                continue
            sel.append(e)
            normals.append(n)
            areas.append(0.5 * np.linalg.norm(nvec))
        if not sel:
            continue
        sel = np.array(sel)
        xe = mesh.node_coords(1)[sel]
        attr = np.arange(next_attr, next_attr + sel.size, dtype=np.int32)
        next_attr += sel.size
        qd = ts.geom_qdata(xe, attr, 1, rpts, np.ones(len(w2)))
        qd[:, 1, :] = w2[None, :] * (2.0 * np.array(areas))[:, None]      # |J e_s x J e_t| = 2 area on a planar face
        groups.append(TetBdrGroup(f, sel, np.ascontiguousarray(interp), qd, np.ascontiguousarray(nd.idx[sel]),
                                  np.ascontiguousarray(nd.curl_orient[sel]), np.array(normals), np.array(areas)))
    return groups


def tangential_materials(groups, c=1.0):
    """Per-face 3 x 3 coefficient c (I - n n^T), in the order of the face attributes: (attr_mat, mat_coeff) for coeff.coeff_ctx."""
    mats = [c * (np.eye(3) - np.outer(n, n)) for g in groups for n in g.normals]
    return np.arange(len(mats)), np.array(mats)
