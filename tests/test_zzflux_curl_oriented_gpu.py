"""GPU: the flux-estimator operators with the tridiagonal ("curl-oriented") element transformations of ND tetrahedra / prisms of
order >= 2 (/root/reference/palace/fem/libceed/restriction.cpp:301-329) on either space: mixed mass operator, table-described
mass operator with its diagonal, projection and element error integrals, against oracle/estimator.py evaluated on element
vectors x_e = T_e x[idx_e] and results scattered through T_e^T. Synthetic small-integer transformations on hexahedral tables (the
kernels are element-agnostic; tests/test_dense_gpu.py does the same for the operator path)."""
import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from oracle import estimator as E
from oracle import pyoracle as O
from palace_b200.host import coeff as cf
from palace_b200.host import hexspace as hs
from tests import common

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def _dev(a):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).cuda()


def _rel(a, b):
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


def tridiag(rng, ne, P):
    co = np.zeros((ne, P, 3), dtype=np.int8)
    co[:, :, 1] = rng.choice([-1, 1], size=(ne, P))
    for e in range(ne):
        for j in rng.choice(np.arange(0, P - 1, 2), size=max(1, P // 6), replace=False):
            b, c = int(rng.integers(-1, 2)), int(rng.integers(-1, 2))
            blk = np.array([[1, b], [0 if b * c == 1 else c, 1]])                   # determinant 1 - b c != 0: invertible
            co[e, j, 1], co[e, j, 2] = blk[0, 0], blk[0, 1]
            co[e, j + 1, 0], co[e, j + 1, 1] = blk[1, 0], blk[1, 1]
    return co


def dense_T(co_e):
    P = co_e.shape[0]
    T = np.zeros((P, P))
    T[np.arange(P), np.arange(P)] = co_e[:, 1]
    T[np.arange(1, P), np.arange(P - 1)] = co_e[1:, 0]
    T[np.arange(P - 1), np.arange(1, P)] = co_e[:-1, 2]
    return T


def assembled(qd, it, mt, idx_t, co_t, nt, is_, ms, idx_s, co_s, ns, coef):
    """sum_e scatter(T_s^T A_e T_t): element matrices from the oracle on a discontinuous numbering, transformed here."""
    ne, Pt, Ps = qd.shape[0], it.shape[2], is_.shape[2]
    dt, ds = np.arange(ne * Pt).reshape(ne, Pt), np.arange(ne * Ps).reshape(ne, Ps)
    Ad = E.mixed_mass_matrix(qd, it, mt, dt, np.ones((ne, Pt)), ne * Pt, is_, ms, ds, np.ones((ne, Ps)), ne * Ps, coef).tocsr()
    rows, cols, vals = [], [], []
    for e in range(ne):
        Ae = Ad[e * Ps:(e + 1) * Ps, e * Pt:(e + 1) * Pt].toarray()
        Me = dense_T(co_s[e]).T @ Ae @ dense_T(co_t[e])
        r, c = np.meshgrid(idx_s[e], idx_t[e], indexing="ij")
        rows.append(r.ravel())
        cols.append(c.ravel())
        vals.append(Me.ravel())
    return sp.csr_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(ns, nt))


def test_estimator_operators_with_curl_oriented_restrictions(b2p_ctx):
    from palace_b200 import capi

    p = 2
    prob = common.make_problem(n=(2, 2, 2), p=p, n_attr=2)
    geom = common.gpu_geom(b2p_ctx, prob)
    nd = prob.nd
    rt = hs.build_rt_space(prob.mesh, prob.topo, p)
    ne = prob.mesh.ne
    rng = np.random.default_rng(12)
    co_n, co_r = tridiag(rng, ne, nd.P), tridiag(rng, ne, rt.P)
    nd_i, rt_i = hs.nd_hex_dense_interp(p, prob.q1d), hs.rt_hex_dense_interp(p, prob.q1d)
    sp_nd = dict(P=nd.P, map_type=capi.MAP_HCURL, interp=nd_i, idx=nd.lex_gid, orient=None, lsize=nd.ndofs, curl_orient=co_n)
    sp_rt = dict(P=rt.P, map_type=capi.MAP_HDIV, interp=rt_i, idx=rt.lex_gid, orient=None, lsize=rt.ndofs, curl_orient=co_r)
    qd = prob.qdata_ref
    muinv = np.array([[[2.0, 0.1, 0.0], [0.1, 1.5, 0.2], [0.0, 0.2, 1.0]], [[1.0, 0.0, 0.3], [0.0, 3.0, 0.0], [0.3, 0.0, 2.0]]])
    c_flux = np.stack([m.ravel(order="F") for m in muinv])
    c_disc = np.stack([E.spd_power(m, 0.5).ravel(order="F") for m in muinv])
    c_smooth = np.stack([E.spd_power(m, -0.5).ravel(order="F") for m in muinv])
    attr = prob.mesh.attr - 1
    celem = lambda tab: [E._mat33(tab[a]) for a in attr]
    I3 = [np.eye(3)] * ne
    # table-described mass operators of both spaces: Mult and diagonal
    for spd, interp, mp, idx, co, n in ((sp_nd, nd_i, E.HCURL, nd.lex_gid, co_n, nd.ndofs), (sp_rt, rt_i, E.HDIV, rt.lex_gid, co_r, rt.ndofs)):
        Mop = capi.vecfe_mass_operator(b2p_ctx, geom, spd)
        Mo = assembled(qd, interp, mp, idx, co, n, interp, mp, idx, co, n, I3)
        x = rng.standard_normal(n)
        y = torch.empty(n, dtype=torch.float64, device="cuda")
        Mop.mult(_dev(x), y)
        assert _rel(y.cpu().numpy(), Mo @ x) < 1e-12
        d = torch.empty(n, dtype=torch.float64, device="cuda")
        Mop.assemble_diagonal(d)
        assert _rel(d.cpu().numpy(), Mo.diagonal()) < 1e-12
    # curl-flux configuration: flux space RT, smooth space ND, both transformed
    Mnd = capi.vecfe_mass_operator(b2p_ctx, geom, sp_nd)
    est = capi.FluxEstimator(b2p_ctx, geom, sp_rt, sp_nd, c_flux, c_disc, c_smooth, Mnd, tol=1e-13, max_it=5000)
    F = assembled(qd, rt_i, E.HDIV, rt.lex_gid, co_r, rt.ndofs, nd_i, E.HCURL, nd.lex_gid, co_n, nd.ndofs, celem(c_flux))
    Mo = assembled(qd, nd_i, E.HCURL, nd.lex_gid, co_n, nd.ndofs, nd_i, E.HCURL, nd.lex_gid, co_n, nd.ndofs, I3)
    B = rng.standard_normal(rt.ndofs)
    H_ref = spla.spsolve(Mo.tocsc(), F @ B)
    Hd = torch.zeros(nd.ndofs, dtype=torch.float64, device="cuda")
    est.project(_dev(B), Hd)
    assert est.stats()["converged"] and _rel(Hd.cpu().numpy(), H_ref) < 1e-9
    # element errors on the transformed element vectors (discontinuous numbering for the oracle)
    Be = np.concatenate([dense_T(co_r[e]) @ B[rt.lex_gid[e]] for e in range(ne)])
    He = np.concatenate([dense_T(co_n[e]) @ H_ref[nd.lex_gid[e]] for e in range(ne)])
    dr, dn = np.arange(ne * rt.P).reshape(ne, rt.P), np.arange(ne * nd.P).reshape(ne, nd.P)
    eta2 = E.element_errors(qd, rt_i, E.HDIV, dr, np.ones((ne, rt.P)), Be, celem(c_disc), nd_i, E.HCURL, dn, np.ones((ne, nd.P)), He,
                            celem(c_smooth))
    ed = torch.zeros(ne, dtype=torch.float64, device="cuda")
    est.integrate(_dev(B), Hd, ed)
    assert _rel(ed.cpu().numpy(), eta2) < 1e-8
