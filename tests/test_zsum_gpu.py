"""GPU parity of the fused REAL sum operator: a ParOperator over several sum-factorised ND terms (a0 K + a1 C + a2 M,
BuildParSumOperator, /root/reference/palace/linalg/rap.cpp:764-829) runs as ONE element operator whose per-element coefficient
tensors are the weighted sums of the terms' (b2p_op_create_sum); against the term-by-term path and the oracle's matrices."""
import numpy as np
import pytest

from oracle import pyoracle as O
from tests import common

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu
RTOL = 1e-12


def _dev(a):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).cuda()


def _rel(a, b):
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


@pytest.mark.parametrize("p", [1, 3, 4])
def test_real_sum_operator_is_fused_and_exact(b2p_ctx, monkeypatch, p):
    from palace_b200 import capi

    prob = common.make_problem(n=(3, 2, 2), p=p, n_attr=3)
    geom = common.gpu_geom(b2p_ctx, prob)
    nd = prob.nd
    specs = [(O.CURLCURL, common.coefficient(O.CURLCURL, 3, "matrix")), (O.ND_MASS, common.coefficient(O.ND_MASS, 3, "matrix", a_mass=1.7)),
             (O.ND_MASS, common.coefficient(O.ND_MASS, 3, "scalar", a_mass=0.3)), (O.CURLCURL_MASS, common.coefficient(O.CURLCURL_MASS, 3, "matrix"))]
    ops = [common.gpu_op(b2p_ctx, geom, prob, k, b) for k, b in specs]
    mats = [common.oracle_matrix(prob, k, b, eliminate=False) for k, b in specs]
    coefs = [1.0, -2.5, 0.8, 0.35]
    monkeypatch.setenv("B2P_SUM_FUSED", "0")
    A0 = capi.Operator.par(b2p_ctx, nd.ndofs, nd.ndofs, ops, coefs, nd.ess_dofs, diag_policy=1)
    monkeypatch.setenv("B2P_SUM_FUSED", "1")
    A1 = capi.Operator.par(b2p_ctx, nd.ndofs, nd.ndofs, ops, coefs, nd.ess_dofs, diag_policy=1)
    assert not A0.is_fused() and A1.is_fused()

    def reference(cf):
        Ao = sum(c * M for c, M in zip(cf, mats)).tolil()
        ess = nd.ess_dofs
        Ao[ess, :] = 0
        Ao[:, ess] = 0
        Ao[ess, ess] = 1.0
        return Ao.tocsr()

    x = np.random.default_rng(4).random(nd.ndofs)
    for cf in (coefs, [0.5, 1.5, -0.25, 2.0]):
        for A in (A0, A1):
            A.set_coefficients(cf)
            y = torch.full((nd.ndofs,), 9.0, dtype=torch.float64, device="cuda")
            A.mult(_dev(x), y)
            assert _rel(y.cpu().numpy(), reference(cf) @ x) < RTOL
            A.add_mult(_dev(x), y, -0.5)
            assert _rel(y.cpu().numpy(), 0.5 * (reference(cf) @ x)) < RTOL
            d = torch.empty(nd.ndofs, dtype=torch.float64, device="cuda")
            A.assemble_diagonal(d)
            assert _rel(d.cpu().numpy(), reference(cf).diagonal()) < RTOL


def test_mixed_element_types_keep_one_apply_per_term(b2p_ctx):
    """A hex volume term plus a boundary (dense) term cannot share one pass: the operator stays term by term and exact."""
    from palace_b200 import capi
    from palace_b200.host import bdrspace as bs
    from palace_b200.host import coeff as cf

    prob = common.make_problem(n=(2, 2, 2), p=2, n_attr=1)
    geom = common.gpu_geom(b2p_ctx, prob)
    nd = prob.nd
    mb = cf.coeff_ctx(a=1.0)
    M = common.gpu_op(b2p_ctx, geom, prob, O.ND_MASS, mb)
    faces = bs.boundary_faces(prob.topo)
    sp = bs.build_nd_bdr_space(nd, faces)
    interp, _ = bs.nd_quad_tables(prob.p, prob.q1d)
    qd = bs.pad32_to_33(bs.bdr_qdata(prob.xe, faces, prob.mesh_order, prob.q1d))
    C = capi.Op.create_dense(b2p_ctx, capi.Geom.general(b2p_ctx, qd), O.ND_MASS, nd.ndofs, sp.idx, sp.orient, interp, None, mb)
    A = capi.Operator.par(b2p_ctx, nd.ndofs, nd.ndofs, [M, C], [1.0, 0.7], None, diag_policy=1)
    assert not A.is_fused()
    x = np.random.default_rng(1).random(nd.ndofs)
    y = torch.empty(nd.ndofs, dtype=torch.float64, device="cuda")
    A.mult(_dev(x), y)
    ref = common.oracle_apply(prob, O.ND_MASS, mb, x) + 0.7 * O.apply_add(O.ND_MASS, interp, None, sp.idx, sp.orient, qd, mb, x, np.zeros(nd.ndofs))
    assert _rel(y.cpu().numpy(), ref) < RTOL


def test_sum_operator_can_be_coarsened(b2p_ctx):
    """The fused sum is an ordinary local operator: p-coarsening it (CeedOperatorCoarsen semantics, fine quadrature and
    coefficients shared) gives the sum of the coarsened terms."""
    from palace_b200 import capi
    from palace_b200.host import hexspace as hs

    prob = common.make_problem(p=3, n_attr=2)
    geom = common.gpu_geom(b2p_ctx, prob)
    specs = [(O.CURLCURL, common.coefficient(O.CURLCURL, 2, "matrix")), (O.ND_MASS, common.coefficient(O.ND_MASS, 2, "matrix", a_mass=1.7))]
    ops = [common.gpu_op(b2p_ctx, geom, prob, k, b) for k, b in specs]
    coefs = [1.0, 7.5]
    fine = capi.Op.create_sum(b2p_ctx, ops, coefs)
    x = np.random.default_rng(8).random(prob.nd.ndofs)
    y = torch.empty(prob.nd.ndofs, dtype=torch.float64, device="cuda")
    fine.apply(_dev(x), y)
    ref = sum(c * common.oracle_apply(prob, k, b, x) for c, (k, b) in zip(coefs, specs))
    assert _rel(y.cpu().numpy(), ref) < RTOL
    pc = 2
    ndc = hs.build_nd_space(prob.mesh, prob.topo, pc)
    t = hs.tables_1d(pc, prob.q1d)
    idx, ori = ndc.native_restriction()
    coarse = fine.coarsen(pc, ndc.ndofs, idx, ori, ndc.dof_map, t.Bo, t.Bc, t.Gc)
    xc = np.random.default_rng(9).random(ndc.ndofs)
    yc = torch.empty(ndc.ndofs, dtype=torch.float64, device="cuda")
    coarse.apply(_dev(xc), yc)
    refc = sum(c * common.oracle_apply(prob, k, b, xc, space=ndc, q1d=prob.q1d) for c, (k, b) in zip(coefs, specs))
    assert _rel(yc.cpu().numpy(), refc) < RTOL
