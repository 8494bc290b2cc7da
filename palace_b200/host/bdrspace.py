"""Caller-side stand-in for BOUNDARY integrators on hexahedral meshes: the quadrilateral faces of the boundary as
2-D Nedelec elements embedded in 3-D, which Palace assembles with the dim = 2, space_dim = 3 QFunctions
(/root/reference/palace/fem/qfunctions/32/geom_32_qf.h, hcurl_32_qf.h; selected in fem/integ/vecfemass.cpp:72-105 by
10 * space_dim + dim = 32) for surface impedance, lumped-port and absorbing-boundary terms
(BilinearForm::PartialAssemble loops over boundary elements the same way, fem/bilinearform.cpp:27-107).

No new device kernel is needed for them: with the 3 x 2 factor adj(J)^T / |J| stored as the first two columns of a
3 x 3 matrix whose third column is zero, and the 2-component reference field padded with a zero third component, the
3-D pointwise map w detJ A^T C A u of the ND mass operator IS f_apply_hcurl_32. This module builds those padded tables
and q-data for the dense-basis operator (b2p_op_create_dense, kind ND_MASS); tests/test_bdr_cpu.py pins the embedding
against the reference's own 32 QFunctions (tests/golden/qf32_golden.npz)."""
from __future__ import annotations

import dataclasses

import numpy as np

from . import hexspace as hs


def geom32_qdata(attr, qw, J):
    """Restatement of f_build_geom_factor_32 (geom_32_qf.h:9-34 with AdjJt32, utils_32_qf.h:22-41): J[6][Q] column-major
    3 x 2 -> qdata[8][Q] = {attr, w |J|, adj(J)^T / |J|}, |J| = sqrt(E G - F^2)."""
    J = np.asarray(J, dtype=np.float64)
    E = J[0] ** 2 + J[1] ** 2 + J[2] ** 2
    G = J[3] ** 2 + J[4] ** 2 + J[5] ** 2
    F = J[0] * J[3] + J[1] * J[4] + J[2] * J[5]
    d = np.sqrt(E * G - F * F)
    adj = np.stack([G * J[0] - F * J[3], G * J[1] - F * J[4], G * J[2] - F * J[5],
                    E * J[3] - F * J[0], E * J[4] - F * J[1], E * J[5] - F * J[2]]) / d
    return np.concatenate([np.asarray(attr, dtype=np.float64)[None], (qw * d)[None], adj / d], axis=0)


def geom31_qdata(attr, qw, J):
    """Restatement of f_build_geom_factor_31 (geom_31_qf.h:9-30 with AdjJt31, utils_31_qf.h:20-31): the tangent J[3][Q] of a line
    element in 3-D -> qdata[5][Q] = {attr, w |J|, adj(J)^T / |J| = J / |J|^2}, |J| the length element."""
    J = np.asarray(J, dtype=np.float64)
    d = np.sqrt(J[0] ** 2 + J[1] ** 2 + J[2] ** 2)
    return np.concatenate([np.asarray(attr, dtype=np.float64)[None], (qw * d)[None], J / d / d], axis=0)


def pad31_to_33(qd5):
    """qdata[..., 5, Q] of a line element -> the [..., 11, Q] layout of the 3-D operators: second and third columns zero. With the
    1-component field padded to (u, 0, 0) the 3-D H(curl) mass map is f_apply_hcurl_31 (hcurl_31_qf.h:12-31), so an edge block
    enters the dense-basis operator like a boundary-face block does."""
    qd5 = np.asarray(qd5)
    out = np.zeros(qd5.shape[:-2] + (11, qd5.shape[-1]))
    out[..., :5, :] = qd5
    return out


def pad32_to_33(qd8):
    """qdata[..., 8, Q] of a boundary element -> the [..., 11, Q] layout of the 3-D operators: third column zero."""
    qd8 = np.asarray(qd8)
    out = np.zeros(qd8.shape[:-2] + (11, qd8.shape[-1]))
    out[..., :8, :] = qd8
    return out


@dataclasses.dataclass
class BdrSpace:
    p: int
    P: int                 # 2 p (p + 1) dofs per face
    faces: np.ndarray      # [nb][3] (element, normal axis, side)
    idx: np.ndarray        # [nb][P] int32 global ND dof
    orient: np.ndarray     # [nb][P] int8 +1 / -1


def boundary_faces(topo: hs.HexTopology, select=None) -> np.ndarray:
    ne = topo.face_id.shape[0]
    out = []
    for e in range(ne):
        for nax in range(3):
            for side in range(2):
                if topo.face_nelem[topo.face_id[e, nax, side]] == 1 and (select is None or select(e, nax, side)):
                    out.append((e, nax, side))
    return np.array(out, dtype=np.int64).reshape(-1, 3)


def build_nd_bdr_space(nd: hs.HexSpace, faces: np.ndarray) -> BdrSpace:
    """Face-local dofs: tangential component t1 first (open index along t1 fastest, closed index along t2), then t2
    (closed along t1 fastest, open along t2); each is the element's lexicographic dof with the normal (closed) index on
    the face, so its global number and sign come from the volume space."""
    p = nd.p
    lay = hs._nd_lex_layout(p)
    pos = {t: l for l, t in enumerate(lay)}
    P2 = 2 * p * (p + 1)
    nb = faces.shape[0]
    idx = np.zeros((nb, P2), dtype=np.int32)
    ori = np.zeros((nb, P2), dtype=np.int8)
    for f, (e, nax, side) in enumerate(faces):
        t1, t2 = hs._others(nax)
        o = 0
        for comp, (na, nbb) in ((t1, (p, p + 1)), (t2, (p + 1, p))):
            for m in range(nbb):
                for i in range(na):
                    ix = [0, 0, 0]
                    ix[nax], ix[t1], ix[t2] = side * p, i, m
                    l = pos[(comp, ix[0], ix[1], ix[2])]
                    idx[f, o] = nd.lex_gid[e, l]
                    ori[f, o] = nd.lex_sign[e, l]
                    o += 1
        assert o == P2
    return BdrSpace(p, P2, faces, idx, ori)


def nd_quad_tables(p: int, q1d: int | None = None):
    """interp[3][Q2][P2] of the quadrilateral Nedelec element in the face's (t1, t2) coordinates, zero third component
    (see module docstring); points t1-fastest; (qw2[Q2]) tensor Gauss-Legendre weights."""
    t = hs.tables_1d(p, q1d)
    q = t.Bo.shape[0]
    Bo, Bc = t.Bo, t.Bc                      # [q][p], [q][p+1]
    P2, Q2 = 2 * p * (p + 1), q * q
    interp = np.zeros((3, Q2, P2))
    for qb in range(q):
        for qa in range(q):
            iq = qa + q * qb
            o = 0
            for m in range(p + 1):
                for i in range(p):
                    interp[0, iq, o] = Bo[qa, i] * Bc[qb, m]
                    o += 1
            for m in range(p):
                for i in range(p + 1):
                    interp[1, iq, o] = Bc[qa, i] * Bo[qb, m]
                    o += 1
    qw2 = np.outer(t.qw, t.qw).ravel()        # [qb][qa] -> qa fastest
    return interp, qw2


def bdr_qdata(xe: np.ndarray, faces: np.ndarray, mesh_order: int, q1d: int, attr=None):
    """q-data [nb][8][Q2] of the boundary faces from the volume elements' nodes xe[ne][3][(k+1)^3] (lexicographic
    Gauss-Lobatto nodes): J = [dx/dxi_t1, dx/dxi_t2] on the face, at tensor Gauss-Legendre points (t1 fastest)."""
    nodes = hs.gauss_lobatto(mesh_order + 1)
    qx, qw = hs.gauss_legendre(q1d)
    B, G = hs.lagrange_table(nodes, qx)                       # [q][k+1]
    Bs, Gs = hs.lagrange_table(nodes, np.array([0.0, 1.0]))    # values on the two sides
    n = mesh_order + 1
    nb, Q2 = faces.shape[0], q1d * q1d
    qw2 = np.outer(qw, qw).ravel()
    out = np.zeros((nb, 8, Q2))
    for f, (e, nax, side) in enumerate(faces):
        t1, t2 = hs._others(nax)
        X = xe[e].reshape(3, n, n, n)                         # [c][k][j][i] (i fastest = axis 0)
        X = np.transpose(X, (0, 3, 2, 1))                     # [c][i][j][k] indexed by axis 0, 1, 2
        J = np.zeros((6, Q2))
        for qb in range(q1d):
            for qa in range(q1d):
                w = [None, None, None]
                w[nax] = Bs[side]
                for d, tdir in ((0, t1), (1, t2)):
                    for tt, qq in ((t1, qa), (t2, qb)):
                        w[tt] = G[qq] if tt == tdir else B[qq]
                    J[3 * d:3 * d + 3, qa + q1d * qb] = np.einsum("cijk,i,j,k->c", X, w[0], w[1], w[2])
        a = np.full(Q2, 1.0 if attr is None else float(attr[f]))
        out[f] = geom32_qdata(a, qw2, J)
    return out


# ---- boundary curl-curl: CurlCurlIntegrator on boundary elements (second-order absorbing boundaries, wave ports;
# /root/reference/palace/models/spaceoperator.cpp:290-300). For dim = 2 in space_dim = 3 the curl has ONE component and the
# integrator selects f_apply_l2_1 with the quadrature weight as an extra input (integ/curlcurl.cpp:54-68):
#     v = coeff qw^2 / (w |J|) curl^ u        (qfunctions/1/l2_1_qf.h:9-22; coefficient context of dimension 1)
# since the reference-space scalar curl is |J| times the physical surface curl. Again no new device kernel: the 3-D curl-curl
# map w detJ Jd^T C Jd c with the geometry factor replaced by the identity (Jd = cofactor(I) = I), the weight slot holding
# qw^2 / (w |J|), the scalar curl padded to (curl, 0, 0) and the scalar coefficient as a diagonal 3 x 3 IS that QFunction. ----
def nd_quad_curl_tables(p: int, q1d: int | None = None):
    """deriv[3][Q2][P2]: component 0 = reference scalar curl d(u_t2)/d(t1) - d(u_t1)/d(t2) of the quadrilateral Nedelec element
    (dof order of build_nd_bdr_space / nd_quad_tables), components 1, 2 zero."""
    t = hs.tables_1d(p, q1d)
    q = t.Bo.shape[0]
    P2, Q2 = 2 * p * (p + 1), q * q
    deriv = np.zeros((3, Q2, P2))
    for qb in range(q):
        for qa in range(q):
            iq = qa + q * qb
            o = 0
            for m in range(p + 1):          # u_t1 = Bo_i(a) Bc_m(b)
                for i in range(p):
                    deriv[0, iq, o] = -t.Bo[qa, i] * t.Gc[qb, m]
                    o += 1
            for m in range(p):              # u_t2 = Bc_i(a) Bo_m(b)
                for i in range(p + 1):
                    deriv[0, iq, o] = t.Gc[qa, i] * t.Bo[qb, m]
                    o += 1
    return deriv


def curl32_qdata(qd8, qw2):
    """[..., 8, Q] boundary q-data + the rule's weights -> the [..., 11, Q] layout the 3-D curl-curl operator reads:
    {attr, qw^2 / (w |J|), identity}."""
    qd8 = np.asarray(qd8)
    out = np.zeros(qd8.shape[:-2] + (11, qd8.shape[-1]))
    out[..., 0, :] = qd8[..., 0, :]
    out[..., 1, :] = np.asarray(qw2) ** 2 / qd8[..., 1, :]
    out[..., 2, :] = out[..., 6, :] = out[..., 10, :] = 1.0
    return out
