"""GPU parity of the boundary curl-curl integrator (CurlCurlIntegrator on boundary elements: second-order absorbing boundaries,
wave ports; /root/reference/palace/models/spaceoperator.cpp:290-300, fem/integ/curlcurl.cpp case 32 -> f_apply_l2_1) through the
dense-basis operator on the padded scalar-curl table and the identity-geometry q-data of palace_b200/host/bdrspace.py (the
embedding is pinned to the reference's l2_1_qf.h in tests/test_bdr_cpu.py), alone and together with the boundary mass as the
CurlCurlMassIntegrator pair of one absorbing face set."""
import numpy as np
import pytest

from oracle import pyoracle as O
from palace_b200.host import bdrspace as bs
from palace_b200.host import coeff as cf
from tests import common

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu
RTOL = 1e-12


def _dev(a):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).cuda()


def _rel(a, b):
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


@pytest.mark.parametrize("p", [1, 2, 3])
def test_boundary_curlcurl_apply_and_diagonal_match_oracle(b2p_ctx, p):
    from palace_b200 import capi

    prob = common.make_problem(n=(3, 2, 2), p=p, n_attr=2)            # curved (order-2, warped) hexes, scrambled frames
    faces = bs.boundary_faces(prob.topo)
    sp = bs.build_nd_bdr_space(prob.nd, faces)
    deriv = bs.nd_quad_curl_tables(prob.p, prob.q1d)
    _, qw2 = bs.nd_quad_tables(prob.p, prob.q1d)
    qd8 = bs.bdr_qdata(prob.xe, faces, prob.mesh_order, prob.q1d, attr=1 + np.arange(faces.shape[0]) % 2)
    qd = bs.curl32_qdata(qd8, qw2)
    blob = cf.widen_scalar_ctx(cf.coeff_ctx(np.array([0, 1]), np.array([1.7, 0.4]), a=1.0, dim=1))
    n = prob.nd.ndofs
    geom = capi.Geom.general(b2p_ctx, qd)
    op = capi.Op.create_dense(b2p_ctx, geom, O.CURLCURL, n, sp.idx, sp.orient, None, deriv, blob)
    x = np.random.default_rng(0).random(n) - 0.5
    y_ref = O.apply_add(O.CURLCURL, None, deriv, sp.idx, sp.orient, qd, blob, x, np.zeros(n))
    y = torch.full((n,), -3.0, dtype=torch.float64, device="cuda")
    op.apply(_dev(x), y)
    assert _rel(y.cpu().numpy(), y_ref) < RTOL
    d = torch.zeros(n, dtype=torch.float64, device="cuda")
    op.diag_add(d)
    d_ref = O.diag_add(O.CURLCURL, None, deriv, sp.idx, qd, blob, np.zeros(n))
    assert _rel(d.cpu().numpy(), d_ref) < RTOL
    # the surface curl-curl form is positive semi-definite and annihilates gradients' tangential traces only through the faces
    assert x @ y_ref > 0


def test_boundary_curlcurl_and_mass_as_two_terms_of_one_operator(b2p_ctx):
    """a.AddBoundaryIntegrator<CurlCurlMassIntegrator>(dfb, fb): the two boundary forms on one face set as two terms of a
    ParOperator next to nothing else (no essential dofs), against the sum of the oracle's applies."""
    from palace_b200 import capi

    prob = common.make_problem(n=(2, 2, 2), p=2, n_attr=1)
    faces = bs.boundary_faces(prob.topo)
    sp = bs.build_nd_bdr_space(prob.nd, faces)
    interp, qw2 = bs.nd_quad_tables(prob.p, prob.q1d)
    deriv = bs.nd_quad_curl_tables(prob.p, prob.q1d)
    qd8 = bs.bdr_qdata(prob.xe, faces, prob.mesh_order, prob.q1d)
    qm, qc = bs.pad32_to_33(qd8), bs.curl32_qdata(qd8, qw2)
    bm = cf.coeff_ctx(a=0.6)
    bc = cf.widen_scalar_ctx(cf.coeff_ctx(np.array([0]), np.array([2.5]), a=1.0, dim=1))
    n = prob.nd.ndofs
    om = capi.Op.create_dense(b2p_ctx, capi.Geom.general(b2p_ctx, qm), O.ND_MASS, n, sp.idx, sp.orient, interp, None, bm)
    oc = capi.Op.create_dense(b2p_ctx, capi.Geom.general(b2p_ctx, qc), O.CURLCURL, n, sp.idx, sp.orient, None, deriv, bc)
    A = capi.Operator.par(b2p_ctx, n, n, [oc, om], [1.0, -0.3], None, diag_policy=1)
    x = np.random.default_rng(1).random(n) - 0.5
    y_ref = (O.apply_add(O.CURLCURL, None, deriv, sp.idx, sp.orient, qc, bc, x, np.zeros(n))
             - 0.3 * O.apply_add(O.ND_MASS, interp, None, sp.idx, sp.orient, qm, bm, x, np.zeros(n)))
    y = torch.empty(n, dtype=torch.float64, device="cuda")
    A.mult(_dev(x), y)
    assert _rel(y.cpu().numpy(), y_ref) < RTOL
