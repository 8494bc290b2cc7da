"""GPU: boundary mass on tetrahedra (surface impedance / lumped-port resistive terms; spaceoperator.cpp:300-303) through the dense-
basis operator fed with the parent tetrahedron's tables at the face points (palace_b200/host/tetbdr.py, checked against exact
surface integrals in tests/test_tetbdr_cpu.py): every face group against the oracle, and K - w^2 M + i w C_bdr with the impedance
faces on one side of the box as ONE complex operator over dense terms of two kinds (domain terms fused pairwise are not: the
boundary groups live on their own geometry handles)."""
import numpy as np
import pytest

from oracle import pyoracle as O
from palace_b200.host import coeff as cf
from palace_b200.host import tetbdr as tb
from palace_b200.host import tetspace as ts

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu
RTOL = 1e-12


def _dev(a):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).cuda()


def _rel(a, b):
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


@pytest.mark.parametrize("p", [1, 2, 3])
def test_tet_boundary_mass_matches_oracle(b2p_ctx, p):
    from palace_b200 import capi

    mesh = ts.box_tet_mesh((3, 2, 2), (1.0, 0.8, 0.9), jitter=0.2, scramble_seed=5)
    nd = ts.build_nd_tet_space(mesh, p)
    groups = tb.boundary_groups(mesh, nd)
    am, mats = tb.tangential_materials(groups, c=0.8)
    blob = cf.coeff_ctx(am, mats)
    x = np.random.default_rng(0).random(nd.ndofs) - 0.5
    y_ref = np.zeros(nd.ndofs)
    ops = []
    for g in groups:
        O.apply_add_co(O.ND_MASS, g.interp, None, g.idx, g.curl_orient, g.qdata, blob, x, y_ref)
        geom = capi.Geom.general(b2p_ctx, g.qdata)
        ops.append(capi.Op.create_dense(b2p_ctx, geom, O.ND_MASS, nd.ndofs, g.idx, None, g.interp, None, blob, curl_orient=g.curl_orient))
    A = capi.Operator.par(b2p_ctx, nd.ndofs, nd.ndofs, ops, None, None, diag_policy=1)
    assert not A.is_fused()
    y = torch.empty(nd.ndofs, dtype=torch.float64, device="cuda")
    A.mult(_dev(x), y)
    assert _rel(y.cpu().numpy(), y_ref) < RTOL


def test_impedance_boundary_on_one_side_of_a_tet_box(b2p_ctx):
    from palace_b200 import capi

    p = 2
    size = (1.0, 0.8, 0.9)
    mesh = ts.box_tet_mesh((2, 2, 2), size, jitter=0.2, scramble_seed=7)
    nd = ts.build_nd_tet_space(mesh, p)
    interp, curl, qpts, qw = ts.nd_tet_tables(p)
    qd = ts.geom_qdata(mesh.node_coords(1), mesh.attr, 1, qpts, qw)
    geom = capi.Geom.general(b2p_ctx, qd)
    one = cf.coeff_ctx(a=1.0)
    K = capi.Op.create_dense(b2p_ctx, geom, O.CURLCURL, nd.ndofs, nd.idx, None, None, curl, one, curl_orient=nd.curl_orient)
    M = capi.Op.create_dense(b2p_ctx, geom, O.ND_MASS, nd.ndofs, nd.idx, None, interp, None, one, curl_orient=nd.curl_orient)
    groups = tb.boundary_groups(mesh, nd, select=lambda c, n: abs(c[0] - size[0]) < 1e-9)     # the x = max side
    assert abs(sum(g.areas.sum() for g in groups) - size[1] * size[2]) < 1e-12
    am, mats = tb.tangential_materials(groups, c=1.0)
    blob = cf.coeff_ctx(am, mats)
    bops = [capi.Op.create_dense(b2p_ctx, capi.Geom.general(b2p_ctx, g.qdata), O.ND_MASS, nd.ndofs, g.idx, None, g.interp, None, blob,
                                 curl_orient=g.curl_orient) for g in groups]
    # PEC on the other five sides: essential dofs = boundary dofs not on the impedance side
    probe = np.zeros(nd.ndofs)
    for j in range(8):   # dofs the impedance form touches (random probes: a dof with a tangential trace on the side gets a response)
        r = np.random.default_rng(j).random(nd.ndofs)
        for g in groups:
            O.apply_add_co(O.ND_MASS, g.interp, None, g.idx, g.curl_orient, g.qdata, blob, r, probe)
    on_side = np.abs(probe) > 1e-12 * np.abs(probe).max()
    ess = np.array([d for d in nd.ess_dofs if not on_side[d]])
    w = 1.7
    coefs = [1.0, -w * w] + [1j * w * 0.6] * len(bops)
    Z = capi.ComplexOperator.par(b2p_ctx, nd.ndofs, nd.ndofs, [K, M] + bops, coefs, ess, diag_policy=1)
    rng = np.random.default_rng(3)
    x = rng.random(nd.ndofs) - 0.5 + 1j * (rng.random(nd.ndofs) - 0.5)
    yr, yi = torch.empty(nd.ndofs, dtype=torch.float64, device="cuda"), torch.empty(nd.ndofs, dtype=torch.float64, device="cuda")
    Z.mult(_dev(x.real), _dev(x.imag), yr, yi)
    # oracle: the same sum term by term with masked input / overwritten rows (ParOperator semantics)
    xm = x.copy()
    xm[ess] = 0.0
    ref = np.zeros(nd.ndofs, dtype=complex)
    for part in (0, 1):
        v = np.ascontiguousarray(xm.real if part == 0 else xm.imag)
        acc = np.zeros(nd.ndofs, dtype=complex)
        t = np.zeros(nd.ndofs)
        O.apply_add_co(O.CURLCURL, interp, curl, nd.idx, nd.curl_orient, qd, one, v, t)
        acc += coefs[0] * t
        t = np.zeros(nd.ndofs)
        O.apply_add_co(O.ND_MASS, interp, curl, nd.idx, nd.curl_orient, qd, one, v, t)
        acc += coefs[1] * t
        t = np.zeros(nd.ndofs)
        for g in groups:
            O.apply_add_co(O.ND_MASS, g.interp, None, g.idx, g.curl_orient, g.qdata, blob, v, t)
        acc += coefs[2] * t
        ref += acc if part == 0 else 1j * acc
    ref[ess] = x[ess]
    assert _rel(yr.cpu().numpy() + 1j * yi.cpu().numpy(), ref) < RTOL
