"""GPU parity of boundary integrators (surface impedance / lumped port / absorbing terms: ND mass on the quadrilateral
boundary faces, the reference's dim = 2 in space_dim = 3 QFunctions, /root/reference/palace/fem/qfunctions/32) through the
dense-basis operator on zero-padded tables and q-data (palace_b200/host/bdrspace.py), alone and as the imaginary term of
a driven-type complex system  K - w^2 M + i w C_bdr  over split real/imag vectors."""
import numpy as np
import pytest

from oracle import pyoracle as O
from oracle import solvers as S
from palace_b200.host import bdrspace as bs
from palace_b200.host import coeff as cf
from palace_b200.host import hexspace as hs
from tests import common

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu
RTOL = 1e-12


def _dev(a):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).cuda()


def _rel(a, b):
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


def _bdr(prob, select=None):
    faces = bs.boundary_faces(prob.topo, select)
    sp = bs.build_nd_bdr_space(prob.nd, faces)
    interp, _ = bs.nd_quad_tables(prob.p, prob.q1d)
    qd = bs.pad32_to_33(bs.bdr_qdata(prob.xe, faces, prob.mesh_order, prob.q1d))
    return sp, interp, qd


@pytest.mark.parametrize("p", [1, 2, 3])
def test_boundary_mass_apply_matches_oracle(b2p_ctx, p):
    from palace_b200 import capi

    prob = common.make_problem(n=(3, 2, 2), p=p, n_attr=1)           # curved (order-2, warped) hexes, scrambled frames
    sp, interp, qd = _bdr(prob)
    blob = cf.coeff_ctx(a=1.7)
    geom = capi.Geom.general(b2p_ctx, qd)
    op = capi.Op.create_dense(b2p_ctx, geom, O.ND_MASS, prob.nd.ndofs, sp.idx, sp.orient, interp, None, blob)
    x = np.random.default_rng(0).random(prob.nd.ndofs)
    y_ref = O.apply_add(O.ND_MASS, interp, None, sp.idx, sp.orient, qd, blob, x, np.zeros(prob.nd.ndofs))
    y = torch.full((prob.nd.ndofs,), -3.0, dtype=torch.float64, device="cuda")
    op.apply(_dev(x), y)
    assert _rel(y.cpu().numpy(), y_ref) < RTOL
    d = torch.zeros(prob.nd.ndofs, dtype=torch.float64, device="cuda")
    op.diag_add(d)
    d_ref = O.diag_add(O.ND_MASS, interp, None, sp.idx, qd, blob, np.zeros(prob.nd.ndofs))
    assert _rel(d.cpu().numpy(), d_ref) < RTOL


def test_impedance_boundary_in_a_complex_system(b2p_ctx):
    """A = K - w^2 M + i w C_bdr with the boundary term on one side of the box only (the other faces PEC)."""
    from palace_b200 import capi

    prob = common.make_problem(n=(3, 2, 2), p=2, n_attr=2)
    nd = prob.nd
    geom = common.gpu_geom(b2p_ctx, prob)
    kb, mb = cf.coeff_ctx(a=1.0), common.coefficient(O.ND_MASS, 2, "matrix")
    K, M = common.gpu_op(b2p_ctx, geom, prob, O.CURLCURL, kb), common.gpu_op(b2p_ctx, geom, prob, O.ND_MASS, mb)
    # impedance faces: boundary faces whose centre lies on the x = max side of the box
    xmax = prob.xe[:, 0, :].max()
    faces_all = bs.boundary_faces(prob.topo)
    # pick by the face's own node coordinates: all 4 corner x-coordinates near xmax
    n1 = prob.mesh_order + 1
    keep = []
    for (e, nax, side) in faces_all:
        X = np.transpose(prob.xe[e].reshape(3, n1, n1, n1), (0, 3, 2, 1))
        sl = [slice(None)] * 3
        sl[nax] = side * (n1 - 1)
        if np.abs(X[0][tuple(sl)] - xmax).max() < 0.12:
            keep.append((e, nax, side))
    keep = set(keep)
    sp, interp, qd = _bdr(prob, lambda e, nax, side: (e, nax, side) in keep)
    assert 0 < sp.faces.shape[0] < faces_all.shape[0]
    cb = cf.coeff_ctx(a=0.8)
    C = capi.Op.create_dense(b2p_ctx, capi.Geom.general(b2p_ctx, qd), O.ND_MASS, nd.ndofs, sp.idx, sp.orient, interp, None, cb)
    # essential dofs: the PEC part of the boundary = boundary dofs not touched by the impedance faces
    ess = np.setdiff1d(nd.ess_dofs, np.unique(sp.idx))
    w = 2.3
    coefs = [1.0 + 0.0j, -w * w + 0.0j, 1j * w]
    A = capi.ComplexOperator.par(b2p_ctx, nd.ndofs, nd.ndofs, [K, M, C], coefs, ess, 1)
    Ko = common.oracle_matrix(prob, O.CURLCURL, kb, eliminate=False)
    Mo = common.oracle_matrix(prob, O.ND_MASS, mb, eliminate=False)
    Ce = O.element_matrices(O.ND_MASS, interp, None, sp.orient, qd, cb, sp.P)
    Co = S.assemble_sparse(Ce, sp.idx.astype(np.int64), nd.ndofs)
    Ao = (coefs[0] * Ko + coefs[1] * Mo + coefs[2] * Co).tolil()
    Ao[ess, :] = 0
    Ao[:, ess] = 0
    Ao[ess, ess] = 1.0
    rng = np.random.default_rng(5)
    x = rng.random(nd.ndofs) + 1j * rng.random(nd.ndofs)
    yr, yi = torch.empty(nd.ndofs, dtype=torch.float64, device="cuda"), torch.empty(nd.ndofs, dtype=torch.float64, device="cuda")
    A.mult(_dev(x.real), _dev(x.imag), yr, yi)
    assert _rel(yr.cpu().numpy() + 1j * yi.cpu().numpy(), Ao.tocsr() @ x) < RTOL
