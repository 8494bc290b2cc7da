"""CPU checks of the tetrahedral Nedelec path (caller-side stand-in palace_b200/host/tetspace.py + the
oracle's dense curl-oriented apply): what Palace feeds libCEED for simplices
(/root/reference/palace/fem/libceed/basis.cpp:40-85, restriction.cpp:281-368). MFEM is not available
here, so the element, the face-pair transformations and the q-data layout are pinned by identities that
fail for any wrong node, tangent, sign, permutation or Piola convention."""
import numpy as np
import pytest

from oracle import pyoracle as O
from palace_b200.host import coeff as cf
from palace_b200.host import tetspace as ts


def _problem(p, n=(2, 1, 1), jitter=0.25, seed=5, geom_order=1, warp=0.0, degree=None, n_attr=1):
    mesh = ts.box_tet_mesh(n, (1.0, 0.8, 0.9), jitter=jitter, scramble_seed=seed, n_attr=n_attr, warp_amp=warp)
    sp = ts.build_nd_tet_space(mesh, p)
    interp, curl, qpts, qw = ts.nd_tet_tables(p, degree)
    qd = ts.geom_qdata(mesh.node_coords(geom_order), mesh.attr, geom_order, qpts, qw)
    return mesh, sp, interp, curl, qpts, qw, qd


def _apply(kind, sp, interp, curl, qd, blob, x):
    return O.apply_add_co(kind, interp, curl, sp.idx, sp.curl_orient, qd, blob, np.ascontiguousarray(x), np.zeros(sp.ndofs))


@pytest.mark.parametrize("p", [1, 2, 3, 4, 5, 6])
def test_reference_element_is_unisolvent_and_dual(p):
    el = ts.nd_tet_element(p)
    assert el.P == ts.nd_tet_ndof(p) == p * (p + 2) * (p + 3) // 2
    assert el.cond < 1e7
    interp, _ = el.tabulate(el.nodes)                      # [3][P nodes][P shapes]
    dual = np.einsum("cnd,nc->nd", interp, el.tangents)    # functional n of shape d
    assert np.abs(dual - np.eye(el.P)).max() < 1e-9 * max(1.0, el.cond * 1e-4)


def test_quadrature_is_exact_to_its_degree():
    from math import factorial

    for deg in (2, 5, 8, 12):
        pts, w = ts.tet_quadrature(deg)
        assert (pts >= 0).all() and (pts.sum(axis=1) <= 1 + 1e-14).all() and (w > 0).all()
        for a in range(deg + 1):
            for b in range(deg + 1 - a):
                c = deg - a - b
                exact = factorial(a) * factorial(b) * factorial(c) / factorial(a + b + c + 3)
                assert abs((w * pts[:, 0] ** a * pts[:, 1] ** b * pts[:, 2] ** c).sum() - exact) < 1e-15


def test_symmetric_24_point_rule_is_exact_to_degree_6():
    """The rule MFEM returns for tetrahedra at order 6 (what the reference integrates order-3 forms with)."""
    from math import factorial

    pts, w = ts.tet_quadrature_symmetric6()
    assert len(w) == 24 and (w > 0).all() and (pts > 0).all() and (pts.sum(axis=1) < 1).all()
    for deg in range(7):
        for a in range(deg + 1):
            for b in range(deg + 1 - a):
                c = deg - a - b
                exact = factorial(a) * factorial(b) * factorial(c) / factorial(a + b + c + 3)
                assert abs((w * pts[:, 0] ** a * pts[:, 1] ** b * pts[:, 2] ** c).sum() - exact) < 1e-16
    assert abs((w * pts[:, 0] ** 7).sum() - factorial(7) / factorial(10)) > 1e-7  # and no further


def test_symmetric_43_point_rule_is_exact_to_degree_8():
    """A second degree-8 rule (Keast's orbit structure, one negative weight) beside the conical one."""
    from math import factorial

    pts, w = ts.tet_quadrature_symmetric8()
    assert len(w) == 43 and (w < 0).sum() == 1 and (pts > 0).all() and (pts.sum(axis=1) < 1).all()
    for deg in range(9):
        for a in range(deg + 1):
            for b in range(deg + 1 - a):
                c = deg - a - b
                exact = factorial(a) * factorial(b) * factorial(c) / factorial(a + b + c + 3)
                assert abs((w * pts[:, 0] ** a * pts[:, 1] ** b * pts[:, 2] ** c).sum() - exact) < 1e-16
    assert abs((w * pts[:, 0] ** 9).sum() - factorial(9) / factorial(12)) > 1e-9


def test_qdata_layout_matches_the_hex_oracle_convention():
    """Same J -> {w detJ, adj(J)^T/detJ column-major} packing as orc_geom_hex_qdata (geom_33_qf.h:9-34):
    an affine map applied to a hex and to a tet must give the same per-point factors."""
    A = np.array([[1.3, 0.2, -0.1], [0.1, 0.9, 0.3], [-0.2, 0.05, 1.1]])
    # hex: trilinear nodes of the unit cube mapped by A (lexicographic, component-major)
    from palace_b200.host import hexspace as hs

    nodes = hs.gauss_lobatto(2)
    g = np.array([[x, y, z] for z in nodes for y in nodes for x in nodes])
    xe_hex = (g @ A.T).T[None]
    qh = O.geom_hex_qdata(np.ascontiguousarray(xe_hex), np.array([1], dtype=np.int32), 1, 2)
    mesh = ts.TetMesh(ts._REF_VERTS @ A.T, np.array([[0, 1, 2, 3]]), np.array([1], dtype=np.int32))
    pts, w = ts.tet_quadrature(2)
    qt = ts.geom_qdata(mesh.node_coords(1), mesh.attr, 1, pts, w)
    assert np.allclose(qt[0, 2:, 0], qh[0, 2:, 0], rtol=1e-13, atol=1e-14)
    assert np.isclose(qt[0, 1, :].sum(), np.linalg.det(A) / 6.0, rtol=1e-13)


@pytest.mark.parametrize("p", [1, 2, 3, 4])
def test_space_dimension_and_transformations(p):
    mesh = ts.box_tet_mesh((2, 2, 1), jitter=0.2, scramble_seed=11)
    sp = ts.build_nd_tet_space(mesh, p)
    assert sp.ndofs == p * sp.n_edges + p * (p - 1) * sp.n_faces + p * (p - 1) * (p - 2) // 2 * mesh.ne
    assert sorted(np.unique(sp.idx)) == list(range(sp.ndofs))
    co = sp.curl_orient
    assert co.dtype == np.int8 and np.abs(co).max() <= 1            # the reference casts to int8 (restriction.cpp:318-329)
    for e in range(mesh.ne):
        T = sp.dense_T(e)
        assert abs(round(abs(np.linalg.det(T))) - 1) == 0            # unimodular: a change of basis on the shared face
    if p == 1:
        assert not co[:, :, 0].any() and not co[:, :, 2].any()
    else:
        assert co[:, :, 0].any() or co[:, :, 2].any()                # scrambled orders do need off-diagonal entries


def _poly_field(deg, seed):
    """Random vector polynomial of total degree <= deg, its curl, as callables on [n][3] points."""
    rng = np.random.default_rng(seed)
    expo = [(a, b, c) for a in range(deg + 1) for b in range(deg + 1 - a) for c in range(deg + 1 - a - b)]
    coef = rng.standard_normal((3, len(expo)))

    def mono(X, e, d=None):
        e = list(e)
        k = 1.0
        if d is not None:
            if e[d] == 0:
                return np.zeros(X.shape[0])
            k = e[d]
            e[d] -= 1
        return k * X[:, 0] ** e[0] * X[:, 1] ** e[1] * X[:, 2] ** e[2]

    def field(X):
        X = np.atleast_2d(X)
        return np.stack([sum(coef[c, m] * mono(X, e) for m, e in enumerate(expo)) for c in range(3)], axis=1)

    def dfield(X, c, d):
        return sum(coef[c, m] * mono(X, e, d) for m, e in enumerate(expo))

    def curl(X):
        X = np.atleast_2d(X)
        return np.stack([dfield(X, 2, 1) - dfield(X, 1, 2), dfield(X, 0, 2) - dfield(X, 2, 0), dfield(X, 1, 0) - dfield(X, 0, 1)], axis=1)

    return (lambda x: field(x)[0] if np.ndim(x) == 1 else field(x)), curl


@pytest.mark.parametrize("p", [1, 2, 3, 4])
def test_polynomial_fields_are_reproduced_on_scrambled_tets(p):
    """Interpolate E in P_{p-1}^3 through the GLOBAL functionals, pull the dofs back through idx + the
    tridiagonal transformation, evaluate with the reference tables and the covariant Piola map: must
    return E (and curl E through the contravariant map) at every quadrature point of every element."""
    mesh, sp, interp, curl, qpts, qw, qd = _problem(p, n=(2, 2, 1))
    E, curlE = _poly_field(p - 1, 3)
    x = ts.interpolate(mesh, sp, E)
    for e in range(mesh.ne):
        v = mesh.verts[mesh.elems[e]]
        J = np.stack([v[1] - v[0], v[2] - v[0], v[3] - v[0]], axis=1)
        xe = sp.dense_T(e) @ x[sp.idx[e]]
        X = v[0] + qpts @ J.T
        u_ref = np.einsum("cqd,d->qc", interp, xe)
        c_ref = np.einsum("cqd,d->qc", curl, xe)
        u = u_ref @ np.linalg.inv(J)                  # J^-T u_ref, row-vector form
        c = c_ref @ J.T / np.linalg.det(J)            # J c_ref / detJ
        scale = max(1.0, np.abs(E(X)).max())
        assert np.abs(u - E(X)).max() < 1e-10 * scale
        assert np.abs(c - curlE(X)).max() < 1e-9 * max(1.0, np.abs(curlE(X)).max())


@pytest.mark.parametrize("p", [1, 2, 3])
def test_energies_and_curl_of_gradients(p):
    mesh, sp, interp, curl, qpts, qw, qd = _problem(p, n=(2, 1, 2), n_attr=1)
    one = cf.coeff_ctx(a=1.0)
    E, curlE = _poly_field(p - 1, 7)
    x = ts.interpolate(mesh, sp, E)
    Mx = _apply(O.ND_MASS, sp, interp, curl, qd, one, x)
    Kx = _apply(O.CURLCURL, sp, interp, curl, qd, one, x)
    # direct integration of |E|^2 and |curl E|^2 with the same rule (degree 2p >= 2(p-1): exact)
    m_ref = k_ref = 0.0
    for e in range(mesh.ne):
        v = mesh.verts[mesh.elems[e]]
        J = np.stack([v[1] - v[0], v[2] - v[0], v[3] - v[0]], axis=1)
        X = v[0] + qpts @ J.T
        wq = qw * np.linalg.det(J)
        m_ref += (wq * (E(X) ** 2).sum(axis=1)).sum()
        k_ref += (wq * (curlE(X) ** 2).sum(axis=1)).sum()
    assert abs(x @ Mx - m_ref) < 1e-10 * m_ref
    assert abs(x @ Kx - k_ref) < 1e-9 * max(k_ref, 1e-3 * m_ref)
    # gradient of a degree-p scalar lies in P_{p-1}^3: the curl-curl operator annihilates it
    rng = np.random.default_rng(2)
    a = rng.standard_normal(3)
    if p == 1:
        grad = lambda X: np.broadcast_to(a, np.atleast_2d(X).shape) if np.ndim(X) > 1 else a
    else:
        B = rng.standard_normal((3, 3))
        B = B + B.T  # phi = a.x + x^T B x / 2 -> grad = a + B x  (degree 1 <= p - 1)
        grad = lambda X: a + np.asarray(X) @ B
    g = ts.interpolate(mesh, sp, grad)
    Kg = _apply(O.CURLCURL, sp, interp, curl, qd, one, g)
    Mg = _apply(O.ND_MASS, sp, interp, curl, qd, one, g)
    assert np.linalg.norm(Kg) < 1e-10 * np.linalg.norm(Mg)


@pytest.mark.parametrize("p", [2, 3])
def test_operator_is_symmetric_on_curved_tets_with_matrix_coefficients(p):
    mesh, sp, interp, curl, qpts, qw, qd = _problem(p, n=(2, 1, 1), geom_order=2, warp=0.03, n_attr=3, degree=2 * p + 2)
    assert (qd[:, 1, :] > 0).all()
    am, mc = cf.test_suite_coefficient(3, "matrix")
    blob = cf.coeff_ctx_pair(cf.coeff_ctx(am, mc, a=1.3), cf.coeff_ctx(am, mc[::-1].copy(), a=0.7, transpose=True))
    rng = np.random.default_rng(4)
    x, y = rng.random(sp.ndofs), rng.random(sp.ndofs)
    Ax = _apply(O.CURLCURL_MASS, sp, interp, curl, qd, blob, x)
    Ay = _apply(O.CURLCURL_MASS, sp, interp, curl, qd, blob, y)
    assert abs(y @ Ax - x @ Ay) < 1e-12 * abs(y @ Ax)
    assert x @ Ax > 0


def test_curl_oriented_oracle_equals_element_matrix_assembly():
    mesh, sp, interp, curl, qpts, qw, qd = _problem(2, n=(1, 1, 2))
    blob = cf.coeff_ctx_pair(cf.coeff_ctx(a=1.0), cf.coeff_ctx(a=2.0))
    Ae = O.element_matrices(O.CURLCURL_MASS, interp, curl, None, qd, blob, sp.P)
    x = np.random.default_rng(0).random(sp.ndofs)
    y_ref = np.zeros(sp.ndofs)
    for e in range(mesh.ne):
        T = sp.dense_T(e)
        np.add.at(y_ref, sp.idx[e], T.T @ (Ae[e] @ (T @ x[sp.idx[e]])))
    y = _apply(O.CURLCURL_MASS, sp, interp, curl, qd, blob, x)
    assert np.linalg.norm(y - y_ref) < 1e-13 * np.linalg.norm(y_ref)


def _assemble_dense(kind, mesh, interp, deriv, idx, Tfun, n, P, qd):
    one = cf.coeff_ctx(a=1.0)
    Ae = O.element_matrices(kind, interp, deriv, None, qd, one, P)
    A = np.zeros((n, n))
    for e in range(mesh.ne):
        T = np.eye(P) if Tfun is None else Tfun(e)
        A[np.ix_(idx[e], idx[e])] += T.T @ Ae[e] @ T
    return A


def _tri(co):
    P = co.shape[0]
    T = np.zeros((P, P))
    T[np.arange(P), np.arange(P)] = co[:, 1]
    T[np.arange(1, P), np.arange(P - 1)] = co[1:, 0]
    T[np.arange(P - 1), np.arange(1, P)] = co[:-1, 2]
    return T


@pytest.mark.parametrize("p", [1, 2, 3])
def test_transfer_operators_on_tets(p):
    """Discrete gradient and p-prolongation on scrambled tets, with the range-side (dual) transformation of the
    interpolators: K G = 0, G^T M G = H1 stiffness, P^T K P = K_coarse, P^T M P = M_coarse -- to round-off."""
    mesh = ts.box_tet_mesh((2, 2, 1), (1.0, 0.8, 0.9), jitter=0.25, scramble_seed=5)
    nd = ts.build_nd_tet_space(mesh, p)
    h1 = ts.build_h1_tet_space(mesh, nd, p)
    assert sorted(np.unique(h1.idx)) == list(range(h1.ndofs))
    interp, curl, qpts, qw = ts.nd_tet_tables(p)
    qd = ts.geom_qdata(mesh.node_coords(1), mesh.attr, 1, qpts, qw)
    K = _assemble_dense(O.CURLCURL, mesh, interp, curl, nd.idx, nd.dense_T, nd.ndofs, nd.P, qd)
    M = _assemble_dense(O.ND_MASS, mesh, interp, curl, nd.idx, nd.dense_T, nd.ndofs, nd.P, qd)
    vals, grad = ts.h1_tet_element(p).tabulate(qpts)
    assert np.abs(vals.sum(axis=1) - 1.0).max() < 1e-12                      # partition of unity
    A1 = _assemble_dense(O.H1_DIFFUSION, mesh, None, grad, h1.idx, None, h1.ndofs, h1.P, qd)
    dual = ts.dual_orient(nd)
    Td = lambda e: _tri(dual[e])
    for e in range(mesh.ne):
        assert np.abs(Td(e).T @ nd.dense_T(e) - np.eye(nd.P)).max() == 0      # dual rows are T^-T exactly
    G = ts.global_interp_matrix(ts.tet_discrete_gradient(p), h1.idx, None, nd.idx, Td, h1.ndofs, nd.ndofs).toarray()
    assert np.abs(K @ G).max() < 1e-12 * np.abs(K).max()
    assert np.abs(G.T @ M @ G - A1).max() < 1e-12 * np.abs(A1).max()
    if p > 1:
        ndc = ts.build_nd_tet_space(mesh, p - 1)
        Pg = ts.global_interp_matrix(ts.nd_tet_prolongation(p - 1, p), ndc.idx, ndc.dense_T, nd.idx, Td, ndc.ndofs, nd.ndofs).toarray()
        ic, cc = ts.nd_tet_element(p - 1).tabulate(qpts)
        Kc = _assemble_dense(O.CURLCURL, mesh, ic, cc, ndc.idx, ndc.dense_T, ndc.ndofs, ndc.P, qd)
        Mc = _assemble_dense(O.ND_MASS, mesh, ic, cc, ndc.idx, ndc.dense_T, ndc.ndofs, ndc.P, qd)
        assert np.abs(Pg.T @ K @ Pg - Kc).max() < 1e-12 * np.abs(Kc).max()
        assert np.abs(Pg.T @ M @ Pg - Mc).max() < 1e-12 * np.abs(Mc).max()


def test_lowest_order_raviart_thomas_space_and_discrete_curl():
    """RT_0 on scrambled tets: unit outward reference fluxes, normal continuity through the orientation signs (a constant field has
    one flux per face whichever neighbour computes it), and curl of the order-1 ND interpolant = RT_0 interpolant of the curl."""
    mesh = ts.box_tet_mesh((2, 2, 1), (1.0, 0.8, 0.9), jitter=0.25, scramble_seed=5)
    nd = ts.build_nd_tet_space(mesh, 1)
    rt = ts.build_rt0_tet_space(mesh, nd)
    # reference fluxes: int over face g of phi_f . n = delta_fg
    V = ts._REF_VERTS
    for g in range(4):
        oth = [t for t in range(4) if t != g]
        c = V[oth].mean(axis=0)
        nvec = np.cross(V[oth[1]] - V[oth[0]], V[oth[2]] - V[oth[0]]) / 2.0          # area-weighted normal
        if nvec @ (c - V[g]) < 0:
            nvec = -nvec
        flux = ts.rt0_tet_tables(c[None])[:, 0, :].T @ nvec                          # the normal component is constant on a face
        assert np.abs(flux - np.eye(4)[g]).max() < 1e-14
    # a field with constant curl: E = 0.5 w x r + a  ->  curl E = w; ND dofs = edge circulations, RT dofs = face fluxes of w
    w, a = np.array([0.3, -0.7, 0.5]), np.array([0.2, 0.1, -0.4])
    x = ts.interpolate(mesh, nd, lambda X: 0.5 * np.cross(w, X) + a)
    X = mesh.verts
    flux = np.zeros(rt.ndofs)
    for g, F in nd.faces.items():
        flux[F] = 0.5 * np.cross(X[g[1]] - X[g[0]], X[g[2]] - X[g[0]]) @ w
    C = ts.discrete_curl_p1(nd, rt)
    assert np.abs(C @ x - flux).max() < 1e-13
    # element view: the physical field J phi^ / detJ assembled with the signs reproduces the constant w in every element
    interp = ts.rt0_tet_tables(np.array([[0.25, 0.25, 0.25]]))
    for e in range(mesh.ne):
        Xe = X[mesh.elems[e]]
        J = np.stack([Xe[1] - Xe[0], Xe[2] - Xe[0], Xe[3] - Xe[0]], axis=1)
        u = J @ (interp[:, 0, :] @ (rt.orient[e] * flux[rt.idx[e]])) / np.linalg.det(J)
        assert np.abs(u - w).max() < 1e-12


@pytest.mark.parametrize("p", [1, 2, 3, 4])
def test_raviart_thomas_tetrahedra_and_the_discrete_curl(p):
    """RT_{p-1} on scrambled, jittered tets: unisolvent reference element, normal continuity of the assembled space (a polynomial
    field of the space has ONE set of global dofs whichever tetrahedron evaluates them), and the commuting property
    curl(ND interpolant) = RT interpolant of the curl for fields of the ND space."""
    k = p - 1
    el = ts.rt_tet_element(k)
    assert el.P == ts.rt_tet_ndof(k) and el.cond < 1e6
    mesh = ts.box_tet_mesh((2, 2, 1), (1.0, 0.8, 0.9), jitter=0.25, scramble_seed=5)
    nd = ts.build_nd_tet_space(mesh, p)
    rt = ts.build_rt_tet_space(mesh, nd, k)
    # a field of P_{p-1}^3 (inside RT_k): u = polynomial; global dofs from every element must agree
    rng = np.random.default_rng(p)
    A = rng.standard_normal((3, 3))
    c0 = rng.standard_normal(3)

    def curl_field(X):   # curl of E below when p >= 2: constant; for p = 1 take a constant field
        return np.array([A[2, 1] - A[1, 2], A[0, 2] - A[2, 0], A[1, 0] - A[0, 1]]) if p >= 2 else np.zeros(3)

    def E_field(X):
        return (A @ X + c0) if p >= 2 else c0

    u = (lambda X: A @ X + c0) if k >= 1 else (lambda X: c0)
    vals = np.full(rt.ndofs, np.nan)
    for e in range(mesh.ne):
        Xe = mesh.verts[mesh.elems[e]]
        J = np.stack([Xe[1] - Xe[0], Xe[2] - Xe[0], Xe[3] - Xe[0]], axis=1)
        detJ = np.linalg.det(J)
        for i in range(el.P):
            x = Xe[0] + J @ el.nodes[i]
            uhat = detJ * np.linalg.solve(J, u(x))            # u = J u^ / detJ
            d = rt.orient[e, i] * (uhat @ el.dirs[i])
            g = rt.idx[e, i]
            if np.isnan(vals[g]):
                vals[g] = d
            else:
                assert abs(vals[g] - d) < 1e-11 * (1 + abs(d))
    assert not np.isnan(vals).any()
    # reconstruction inside an element reproduces the field
    pts = np.array([[0.2, 0.3, 0.1], [0.25, 0.25, 0.25]])
    tab = el.tabulate(pts)
    for e in (0, mesh.ne - 1):
        Xe = mesh.verts[mesh.elems[e]]
        J = np.stack([Xe[1] - Xe[0], Xe[2] - Xe[0], Xe[3] - Xe[0]], axis=1)
        for q in range(2):
            uq = J @ (tab[:, q, :] @ (rt.orient[e] * vals[rt.idx[e]])) / np.linalg.det(J)
            assert np.abs(uq - u(Xe[0] + J @ pts[q])).max() < 1e-10
    # commuting diagram on the ND interpolant of E
    x_nd = ts.interpolate(mesh, nd, E_field)
    C = ts.tet_discrete_curl(p)
    for e in range(mesh.ne):
        Xe = mesh.verts[mesh.elems[e]]
        J = np.stack([Xe[1] - Xe[0], Xe[2] - Xe[0], Xe[3] - Xe[0]], axis=1)
        detJ = np.linalg.det(J)
        b_loc = C @ (nd.dense_T(e) @ x_nd[nd.idx[e]])
        for i in range(el.P):
            x = Xe[0] + J @ el.nodes[i]
            want = (detJ * np.linalg.solve(J, curl_field(x))) @ el.dirs[i]
            assert abs(b_loc[i] - want) < 1e-10 * (1 + abs(want))
