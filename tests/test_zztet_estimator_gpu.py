"""GPU: the flux error estimators on TETRAHEDRA at lowest order (the element type of BASELINE configs 3 and 4): order-1 Nedelec
and RT_0 spaces of the host layer (tetspace.build_rt0_tet_space, discrete_curl_p1), curl-flux configuration
(CurlFluxErrorEstimator, /root/reference/palace/linalg/errorestimator.cpp:400-513: mu^-1 B projected onto ND, element integrals) and
grad-flux configuration (:272-398: eps E projected onto RT_0 with the table-described RT mass) through the C ABI against
oracle/estimator.py; a field whose flux lies in the smooth space has no error."""
import numpy as np
import pytest
import scipy.sparse.linalg as spla

from oracle import estimator as E
from oracle import pyoracle as O
from palace_b200.host import coeff as cf
from palace_b200.host import tetspace as ts

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def _dev(a):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).cuda()


def _rel(a, b):
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


def test_flux_estimators_on_lowest_order_tetrahedra(b2p_ctx):
    from palace_b200 import capi

    mesh = ts.box_tet_mesh((2, 2, 2), (1.0, 0.8, 0.9), jitter=0.2, scramble_seed=9, n_attr=2)
    nd = ts.build_nd_tet_space(mesh, 1)
    rt = ts.build_rt0_tet_space(mesh, nd)
    C = ts.discrete_curl_p1(nd, rt)
    nd_interp, _, qpts, qw = ts.nd_tet_tables(1, 4)
    rt_interp = ts.rt0_tet_tables(qpts)
    qd = ts.geom_qdata(mesh.node_coords(1), mesh.attr, 1, qpts, qw)
    ne = mesh.ne
    geom = capi.Geom.general(b2p_ctx, qd)
    nd_sign = nd.orient_signs()
    sp_nd = dict(P=nd.P, map_type=capi.MAP_HCURL, interp=nd_interp, idx=nd.idx, orient=nd_sign, lsize=nd.ndofs)
    sp_rt = dict(P=4, map_type=capi.MAP_HDIV, interp=rt_interp, idx=rt.idx, orient=rt.orient, lsize=rt.ndofs)
    mop = capi.Op.create_dense(b2p_ctx, geom, O.ND_MASS, nd.ndofs, nd.idx, nd_sign, nd_interp, None, cf.coeff_ctx(a=1.0))
    Mnd = capi.Operator.par(b2p_ctx, nd.ndofs, nd.ndofs, [mop], None, None, diag_policy=1)
    # two materials: mu^-1 = diag-dominant SPD matrices per attribute
    muinv = np.array([[[2.0, 0.1, 0.0], [0.1, 1.5, 0.2], [0.0, 0.2, 1.0]], [[1.0, 0.0, 0.3], [0.0, 3.0, 0.0], [0.3, 0.0, 2.0]]])
    c_flux = np.stack([m.ravel(order="F") for m in muinv])
    c_disc = np.stack([E.spd_power(m, 0.5).ravel(order="F") for m in muinv])
    c_smooth = np.stack([E.spd_power(m, -0.5).ravel(order="F") for m in muinv])
    est = capi.FluxEstimator(b2p_ctx, geom, sp_rt, sp_nd, c_flux, c_disc, c_smooth, Mnd, tol=1e-13, max_it=5000)
    attr = mesh.attr - 1
    celem = lambda tab: [E._mat33(tab[a]) for a in attr]
    idx_n, ori_n = nd.idx.astype(np.int64), nd_sign.astype(float)
    idx_r, ori_r = rt.idx.astype(np.int64), rt.orient.astype(float)
    F = E.mixed_mass_matrix(qd, rt_interp, E.HDIV, idx_r, ori_r, rt.ndofs, nd_interp, E.HCURL, idx_n, ori_n, nd.ndofs, celem(c_flux))
    Mo = E.mixed_mass_matrix(qd, nd_interp, E.HCURL, idx_n, ori_n, nd.ndofs, nd_interp, E.HCURL, idx_n, ori_n, nd.ndofs, [np.eye(3)] * ne)
    rng = np.random.default_rng(5)
    B = C @ rng.standard_normal(nd.ndofs)                       # a discrete curl: the estimator's input in the reference
    H_ref = spla.spsolve(Mo.tocsc(), F @ B)
    eta2 = E.element_errors(qd, rt_interp, E.HDIV, idx_r, ori_r, B, celem(c_disc), nd_interp, E.HCURL, idx_n, ori_n, H_ref, celem(c_smooth))
    Hd = torch.zeros(nd.ndofs, dtype=torch.float64, device="cuda")
    est.project(_dev(B), Hd)
    assert est.stats()["converged"] and _rel(Hd.cpu().numpy(), H_ref) < 1e-9
    ed = torch.zeros(ne, dtype=torch.float64, device="cuda")
    est.integrate(_dev(B), Hd, ed)
    assert _rel(ed.cpu().numpy(), eta2) < 1e-9
    # one material, constant B: mu^-1 B is a constant field, which the order-1 ND space holds exactly -> no error
    I9 = np.eye(3).ravel()[None].repeat(2, axis=0)
    est1 = capi.FluxEstimator(b2p_ctx, geom, sp_rt, sp_nd, I9, I9, I9, Mnd, tol=1e-14, max_it=5000)
    w = np.array([0.3, -0.7, 0.5])
    X = mesh.verts
    Bc = np.zeros(rt.ndofs)
    for g, Fi in nd.faces.items():
        Bc[Fi] = 0.5 * np.cross(X[g[1]] - X[g[0]], X[g[2]] - X[g[0]]) @ w
    est1.indicator(_dev(Bc), None, 0.0, ed)
    assert float(ed.abs().max()) < 1e-10 * np.linalg.norm(w)
    # grad-flux roles: eps E (E in ND) projected onto RT_0 with the table-described RT mass
    Mrt = capi.vecfe_mass_operator(b2p_ctx, geom, sp_rt)
    Mrt_o = E.mixed_mass_matrix(qd, rt_interp, E.HDIV, idx_r, ori_r, rt.ndofs, rt_interp, E.HDIV, idx_r, ori_r, rt.ndofs, [np.eye(3)] * ne)
    xr = rng.standard_normal(rt.ndofs)
    yr = torch.empty(rt.ndofs, dtype=torch.float64, device="cuda")
    Mrt.mult(_dev(xr), yr)
    assert _rel(yr.cpu().numpy(), Mrt_o @ xr) < 1e-12
    gest = capi.FluxEstimator(b2p_ctx, geom, sp_nd, sp_rt, c_flux, c_disc, c_smooth, Mrt, tol=1e-13, max_it=5000)
    Ev = rng.standard_normal(nd.ndofs)
    Fg = E.mixed_mass_matrix(qd, nd_interp, E.HCURL, idx_n, ori_n, nd.ndofs, rt_interp, E.HDIV, idx_r, ori_r, rt.ndofs, celem(c_flux))
    D_ref = spla.spsolve(Mrt_o.tocsc(), Fg @ Ev)
    eg = E.element_errors(qd, nd_interp, E.HCURL, idx_n, ori_n, Ev, celem(c_disc), rt_interp, E.HDIV, idx_r, ori_r, D_ref, celem(c_smooth))
    Dd = torch.zeros(rt.ndofs, dtype=torch.float64, device="cuda")
    gest.project(_dev(Ev), Dd)
    assert gest.stats()["converged"] and _rel(Dd.cpu().numpy(), D_ref) < 1e-9
    ed.zero_()
    gest.integrate(_dev(Ev), Dd, ed)
    assert _rel(ed.cpu().numpy(), eg) < 1e-8


@pytest.mark.parametrize("p", [2, 3])
def test_curl_flux_estimator_on_higher_order_tetrahedra(b2p_ctx, p):
    """Order-p Nedelec tetrahedra with their REAL tridiagonal element transformations (scrambled vertex orders) as the smooth space,
    RT_{p-1} as the flux space, B a discrete curl: projection and element errors against the oracle on transformed element vectors;
    a field whose curl flux is a polynomial the ND space holds has no error."""
    from palace_b200 import capi
    from tests.test_zzflux_curl_oriented_gpu import assembled, dense_T

    mesh = ts.box_tet_mesh((2, 1, 1), (1.0, 0.8, 0.9), jitter=0.2, scramble_seed=11)
    nd = ts.build_nd_tet_space(mesh, p)
    rt = ts.build_rt_tet_space(mesh, nd, p - 1)
    nd_i, _, qpts, qw = ts.nd_tet_tables(p, 2 * p + 1)
    rt_i = ts.rt_tet_element(p - 1).tabulate(qpts)
    qd = ts.geom_qdata(mesh.node_coords(1), mesh.attr, 1, qpts, qw)
    ne = mesh.ne
    geom = capi.Geom.general(b2p_ctx, qd)
    sp_nd = dict(P=nd.P, map_type=capi.MAP_HCURL, interp=nd_i, idx=nd.idx, orient=None, lsize=nd.ndofs, curl_orient=nd.curl_orient)
    sp_rt = dict(P=rt.P, map_type=capi.MAP_HDIV, interp=rt_i, idx=rt.idx, orient=rt.orient, lsize=rt.ndofs)
    Mnd = capi.vecfe_mass_operator(b2p_ctx, geom, sp_nd)
    I9 = np.eye(3).ravel()[None]
    est = capi.FluxEstimator(b2p_ctx, geom, sp_rt, sp_nd, I9, I9, I9, Mnd, tol=1e-13, max_it=8000)
    co_r = np.zeros((ne, rt.P, 3), dtype=np.int8)
    co_r[:, :, 1] = rt.orient
    I3 = [np.eye(3)] * ne
    F = assembled(qd, rt_i, E.HDIV, rt.idx, co_r, rt.ndofs, nd_i, E.HCURL, nd.idx, nd.curl_orient, nd.ndofs, I3)
    Mo = assembled(qd, nd_i, E.HCURL, nd.idx, nd.curl_orient, nd.ndofs, nd_i, E.HCURL, nd.idx, nd.curl_orient, nd.ndofs, I3)
    C = ts.tet_discrete_curl(p)
    rng = np.random.default_rng(p)
    x = rng.standard_normal(nd.ndofs)
    B = np.zeros(rt.ndofs)
    for e in range(ne):                                      # global RT dofs of curl x (the same from every element: conformity)
        B[rt.idx[e]] = rt.orient[e] * (C @ (nd.dense_T(e) @ x[nd.idx[e]]))
    H_ref = spla.spsolve(Mo.tocsc(), F @ B)
    Hd = torch.zeros(nd.ndofs, dtype=torch.float64, device="cuda")
    est.project(_dev(B), Hd)
    assert est.stats()["converged"] and _rel(Hd.cpu().numpy(), H_ref) < 1e-8
    Be = np.concatenate([rt.orient[e] * B[rt.idx[e]] for e in range(ne)])
    He = np.concatenate([dense_T(nd.curl_orient[e]) @ H_ref[nd.idx[e]] for e in range(ne)])
    dr, dn = np.arange(ne * rt.P).reshape(ne, rt.P), np.arange(ne * nd.P).reshape(ne, nd.P)
    eta2 = E.element_errors(qd, rt_i, E.HDIV, dr, np.ones((ne, rt.P)), Be, I3, nd_i, E.HCURL, dn, np.ones((ne, nd.P)), He, I3)
    ed = torch.zeros(ne, dtype=torch.float64, device="cuda")
    est.integrate(_dev(B), Hd, ed)
    assert _rel(ed.cpu().numpy(), eta2) < 1e-7
    # E = 0.5 w x r: B = w (constant), held exactly by ND_p -> no error
    w = np.array([0.3, -0.7, 0.5])
    xw = ts.interpolate(mesh, nd, lambda X: 0.5 * np.cross(w, X))
    Bw = np.zeros(rt.ndofs)
    for e in range(ne):
        Bw[rt.idx[e]] = rt.orient[e] * (C @ (nd.dense_T(e) @ xw[nd.idx[e]]))
    est.indicator(_dev(Bw), None, 0.0, ed)
    assert float(ed.abs().max()) < 1e-8 * np.linalg.norm(w)


@pytest.mark.parametrize("p", [1, 2, 3])
def test_discrete_curl_operator_on_tetrahedra(b2p_ctx, p):
    """The reference's Curl operator (SpaceOperator::GetCurlMatrix: a DiscreteLinearOperator ND -> RT, assembled like the gradient and
    the prolongations as an element-dense interpolator, fem/bilinearform.cpp:203-282) through b2p_interp_create_dense: ND side with
    the curl-oriented restriction, RT side with its signs; against the host layer's global curl and the transposed product."""
    from palace_b200 import capi

    mesh = ts.box_tet_mesh((2, 2, 1), (1.0, 0.8, 0.9), jitter=0.2, scramble_seed=13)
    nd = ts.build_nd_tet_space(mesh, p)
    rt = ts.build_rt_tet_space(mesh, nd, p - 1)
    C = ts.tet_discrete_curl(p)
    it = capi.Interp.dense(b2p_ctx, C, nd.idx, nd.ndofs, rt.idx, rt.ndofs, in_curl_orient=nd.curl_orient, out_orient=rt.orient)
    Cop = capi.Operator.interp(b2p_ctx, it)
    rng = np.random.default_rng(p)
    x = rng.standard_normal(nd.ndofs)
    B_ref = np.zeros(rt.ndofs)
    for e in range(mesh.ne):
        B_ref[rt.idx[e]] = rt.orient[e] * (C @ (nd.dense_T(e) @ x[nd.idx[e]]))   # the same value from every element sharing a face
    y = torch.empty(rt.ndofs, dtype=torch.float64, device="cuda")
    Cop.mult(_dev(x), y)
    assert _rel(y.cpu().numpy(), B_ref) < 1e-12
    # curl of a gradient vanishes: compose with the discrete gradient
    h1 = ts.build_h1_tet_space(mesh, nd, p)
    git = capi.Interp.dense(b2p_ctx, ts.tet_discrete_gradient(p), h1.idx, h1.ndofs, nd.idx, nd.ndofs, out_curl_orient=ts.dual_orient(nd))
    G = capi.Operator.interp(b2p_ctx, git)
    phi = rng.standard_normal(h1.ndofs)
    g = torch.empty(nd.ndofs, dtype=torch.float64, device="cuda")
    G.mult(_dev(phi), g)
    Cop.mult(g, y)
    assert float(y.abs().max()) < 1e-11 * float(g.abs().max())


@pytest.mark.parametrize("p", [1, 2])
def test_floquet_correction_solver_is_a_flux_projection(b2p_ctx, p):
    """FloquetCorrSolver (/root/reference/palace/linalg/floquetcorrection.cpp:18-84: y = M_rt^-1 Cross x with Cross the mixed
    ND -> RT mass of the cross-product matrix [k x], PCG + Jacobi) is the flux projection with the ND space as flux space, RT as
    smooth space and [k x] as the flux coefficient: b2p_flux_estimator_project. For a constant field E = a the result is the RT
    interpolant of k x a exactly (constants lie in RT_{p-1} on straight-sided tets)."""
    from palace_b200 import capi

    mesh = ts.box_tet_mesh((2, 2, 1), (1.0, 0.8, 0.9), jitter=0.2, scramble_seed=17)
    nd = ts.build_nd_tet_space(mesh, p)
    rt = ts.build_rt_tet_space(mesh, nd, p - 1)
    nd_i, _, qpts, qw = ts.nd_tet_tables(p, 2 * p)
    rt_el = ts.rt_tet_element(p - 1)
    rt_i = rt_el.tabulate(qpts)
    qd = ts.geom_qdata(mesh.node_coords(1), mesh.attr, 1, qpts, qw)
    geom = capi.Geom.general(b2p_ctx, qd)
    sp_nd = dict(P=nd.P, map_type=capi.MAP_HCURL, interp=nd_i, idx=nd.idx, orient=None, lsize=nd.ndofs, curl_orient=nd.curl_orient)
    sp_rt = dict(P=rt.P, map_type=capi.MAP_HDIV, interp=rt_i, idx=rt.idx, orient=rt.orient, lsize=rt.ndofs)
    k = np.array([0.0, 0.3, 0.4])
    kx = np.array([[0.0, -k[2], k[1]], [k[2], 0.0, -k[0]], [-k[1], k[0], 0.0]])
    Mrt = capi.vecfe_mass_operator(b2p_ctx, geom, sp_rt)
    I9 = np.eye(3).ravel()[None]
    corr = capi.FluxEstimator(b2p_ctx, geom, sp_nd, sp_rt, kx.ravel(order="F")[None], I9, I9, Mrt, tol=1e-13, max_it=5000)
    a = np.array([0.7, -1.1, 0.4])
    x = ts.interpolate(mesh, nd, lambda X: a)
    y = torch.zeros(rt.ndofs, dtype=torch.float64, device="cuda")
    corr.project(_dev(x), y)
    assert corr.stats()["converged"]
    want = np.zeros(rt.ndofs)
    u = np.cross(k, a)
    for e in range(mesh.ne):
        Xe = mesh.verts[mesh.elems[e]]
        J = np.stack([Xe[1] - Xe[0], Xe[2] - Xe[0], Xe[3] - Xe[0]], axis=1)
        uhat = np.linalg.det(J) * np.linalg.solve(J, u)
        want[rt.idx[e]] = rt.orient[e] * (rt_el.dirs @ uhat)
    assert _rel(y.cpu().numpy(), want) < 1e-9
