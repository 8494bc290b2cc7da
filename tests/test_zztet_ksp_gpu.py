"""GPU: the BaseKspSolver composer (b2p_ksp_*) on tetrahedra -- the reference's default configuration (FGMRES, p-multigrid, Chebyshev
+ Hiptmair smoothing, linalg/ksp.cpp:131-239, utils/iodata.cpp:500-545) over dense-basis operators with curl-oriented restrictions,
system matrix a0 K + a2 M given as TWO terms (fused into one dense operator), coarse level either Jacobi-smoothed or the Jacobi-PCG
on the device-assembled matrix. The solution must solve the oracle's assembled system."""
import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from oracle import pyoracle as O
from palace_b200.host import coeff as cf
from palace_b200.host import tetspace as ts

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def _dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.mark.parametrize("coarse_type", [0, 1])
def test_ksp_composer_on_tetrahedra(b2p_ctx, coarse_type):
    from palace_b200 import capi

    capi.set_stream(b2p_ctx)
    mesh = ts.box_tet_mesh((2, 2, 1), (1.0, 0.8, 0.9), jitter=0.2, scramble_seed=4)
    orders = [1, 2]
    spaces = [ts.build_nd_tet_space(mesh, p) for p in orders]
    h1s = [ts.build_h1_tet_space(mesh, s_, s_.p) for s_ in spaces]
    _, _, qpts, qw = ts.nd_tet_tables(orders[-1])
    qd = ts.geom_qdata(mesh.node_coords(1), mesh.attr, 1, qpts, qw)
    geom = capi.Geom.general(b2p_ctx, qd)
    one = cf.coeff_ctx(a=1.0)
    a0, a2 = 1.0, 2.5
    pars, auxs, grads, keep = [], [], [], []
    for s_, h1 in zip(spaces, h1s):
        interp, curl = ts.nd_tet_element(s_.p).tabulate(qpts)
        K = capi.Op.create_dense(b2p_ctx, geom, O.CURLCURL, s_.ndofs, s_.idx, None, None, curl, one, curl_orient=s_.curl_orient)
        M = capi.Op.create_dense(b2p_ctx, geom, O.ND_MASS, s_.ndofs, s_.idx, None, interp, None, one, curl_orient=s_.curl_orient)
        keep += [K, M]
        A = capi.Operator.par(b2p_ctx, s_.ndofs, s_.ndofs, [K, M], [a0, a2], s_.ess_dofs, diag_policy=1)
        assert A.is_fused()
        pars.append(A)
        _, grad = ts.h1_tet_element(h1.p).tabulate(qpts)
        aop = capi.Op.create_dense(b2p_ctx, geom, O.H1_DIFFUSION, h1.ndofs, h1.idx, None, None, grad, cf.coeff_ctx(a=a2))
        keep.append(aop)
        auxs.append(capi.Operator.par(b2p_ctx, h1.ndofs, h1.ndofs, [aop], None, h1.ess_dofs, diag_policy=1))
        git = capi.Interp.dense(b2p_ctx, ts.tet_discrete_gradient(s_.p), h1.idx, h1.ndofs, s_.idx, s_.ndofs, out_curl_orient=ts.dual_orient(s_))
        grads.append(capi.Operator.interp(b2p_ctx, git))
    it = capi.Interp.dense(b2p_ctx, ts.nd_tet_prolongation(1, 2), spaces[0].idx, spaces[0].ndofs, spaces[1].idx, spaces[1].ndofs,
                           in_curl_orient=spaces[0].curl_orient, out_curl_orient=ts.dual_orient(spaces[1]))
    prol = [capi.Operator.interp(b2p_ctx, it)]
    ksp = capi.Ksp(b2p_ctx, 2, prol, grads, tol=1e-10, max_it=80, coarse_type=coarse_type, coarse_tol=1e-6, coarse_max_it=400)
    ksp.set_operators(pars[-1], pars, auxs)
    fine = spaces[-1]
    b = np.random.default_rng(2).random(fine.ndofs)
    b[fine.ess_dofs] = 0.0
    x = torch.zeros(fine.ndofs, dtype=torch.float64, device="cuda")
    ksp.mult(_dev(b), x)
    st = ksp.stats()
    assert st["converged"] and st["num_total_mult"] == 1 and st["num_total_mult_its"] == st["its"], st
    interp, curl = ts.nd_tet_element(fine.p).tabulate(qpts)
    blob = cf.coeff_ctx_pair(cf.coeff_ctx(a=a2), cf.coeff_ctx(a=a0))
    Ae = O.element_matrices(O.CURLCURL_MASS, interp, curl, None, qd, blob, fine.P)
    A = np.zeros((fine.ndofs, fine.ndofs))
    for e in range(mesh.ne):
        T = fine.dense_T(e)
        A[np.ix_(fine.idx[e], fine.idx[e])] += T.T @ Ae[e] @ T
    ess = fine.ess_dofs
    A[ess, :] = 0
    A[:, ess] = 0
    A[ess, ess] = 1.0
    xs = x.cpu().numpy()
    assert np.linalg.norm(b - A @ xs) < 1e-8 * np.linalg.norm(b)


def test_complex_multigrid_on_tetrahedra(b2p_ctx):
    """GeometricMultigridSolver<ComplexOperator> with Hiptmair smoothing (the reference's default for complex systems, PCMatReal =
    false) over dense-basis tetrahedral level operators K + (1 + 0.3 i) M -- complex sums over dense terms, i.e. the two-coefficient-sum
    path -- preconditioning complex FGMRES; the solution must solve the assembled complex system."""
    from palace_b200 import capi

    capi.set_stream(b2p_ctx)
    mesh = ts.box_tet_mesh((2, 2, 1), (1.0, 0.8, 0.9), jitter=0.2, scramble_seed=6)
    orders = [1, 2]
    spaces = [ts.build_nd_tet_space(mesh, p) for p in orders]
    h1s = [ts.build_h1_tet_space(mesh, s_, s_.p) for s_ in spaces]
    _, _, qpts, qw = ts.nd_tet_tables(orders[-1])
    qd = ts.geom_qdata(mesh.node_coords(1), mesh.attr, 1, qpts, qw)
    geom = capi.Geom.general(b2p_ctx, qd)
    one = cf.coeff_ctx(a=1.0)
    cm = 1.0 + 0.3j
    A, AG, G, keep = [], [], [], []
    for s_, h1 in zip(spaces, h1s):
        interp, curl = ts.nd_tet_element(s_.p).tabulate(qpts)
        K = capi.Op.create_dense(b2p_ctx, geom, O.CURLCURL, s_.ndofs, s_.idx, None, None, curl, one, curl_orient=s_.curl_orient)
        M = capi.Op.create_dense(b2p_ctx, geom, O.ND_MASS, s_.ndofs, s_.idx, None, interp, None, one, curl_orient=s_.curl_orient)
        _, grad = ts.h1_tet_element(h1.p).tabulate(qpts)
        D = capi.Op.create_dense(b2p_ctx, geom, O.H1_DIFFUSION, h1.ndofs, h1.idx, None, None, grad, one)
        keep += [K, M, D]
        A.append(capi.ComplexOperator.par(b2p_ctx, s_.ndofs, s_.ndofs, [K, M], [1.0, cm], s_.ess_dofs, 1))
        AG.append(capi.ComplexOperator.par(b2p_ctx, h1.ndofs, h1.ndofs, [D], [cm], h1.ess_dofs, 1))
        git = capi.Interp.dense(b2p_ctx, ts.tet_discrete_gradient(s_.p), h1.idx, h1.ndofs, s_.idx, s_.ndofs, out_curl_orient=ts.dual_orient(s_))
        G.append(capi.Operator.interp(b2p_ctx, git))
    it = capi.Interp.dense(b2p_ctx, ts.nd_tet_prolongation(1, 2), spaces[0].idx, spaces[0].ndofs, spaces[1].idx, spaces[1].ndofs,
                           in_curl_orient=spaces[0].curl_orient, out_curl_orient=ts.dual_orient(spaces[1]))
    P = [capi.Operator.interp(b2p_ctx, it)]
    coarse = capi.ComplexSolver.krylov(b2p_ctx, 1, rel_tol=1e-12, max_it=300, max_dim=300)
    mg = capi.ComplexSolver.gmg(b2p_ctx, coarse, P, G, cycle_it=1, smooth_it=1, cheby_order=4)
    mg.gmg_set_operators(A, AG)
    fine = spaces[-1]
    n = fine.ndofs
    rng = np.random.default_rng(3)
    b = rng.standard_normal(n) + 1j * rng.standard_normal(n)
    b[fine.ess_dofs] = 0.0
    k = capi.ComplexSolver.krylov(b2p_ctx, 2, rel_tol=1e-10, max_it=60, max_dim=60)
    k.set_operator(A[-1])
    k.set_preconditioner(mg)
    zr, zi = torch.zeros(n, dtype=torch.float64, device="cuda"), torch.zeros(n, dtype=torch.float64, device="cuda")
    k.mult(_dev(b.real), _dev(b.imag), zr, zi)
    st = k.stats()
    assert st["converged"] and st["its"] <= 30, st
    assert A[-1].fused_applies() > 0
    interp, curl = ts.nd_tet_element(fine.p).tabulate(qpts)
    Ke = O.element_matrices(O.CURLCURL, interp, curl, None, qd, one, fine.P)
    Me = O.element_matrices(O.ND_MASS, interp, curl, None, qd, one, fine.P)
    Z = np.zeros((n, n), dtype=complex)
    for e in range(mesh.ne):
        T = fine.dense_T(e)
        Z[np.ix_(fine.idx[e], fine.idx[e])] += T.T @ (Ke[e] + cm * Me[e]) @ T
    ess = fine.ess_dofs
    Z[ess, :] = 0
    Z[:, ess] = 0
    Z[ess, ess] = 1.0
    z = zr.cpu().numpy() + 1j * zi.cpu().numpy()
    assert np.linalg.norm(b - Z @ z) < 1e-8 * np.linalg.norm(b)


def test_divfree_projection_on_tetrahedra(b2p_ctx):
    """DivFreeSolver (linalg/divfree.cpp:42-186) over dense tetrahedral operators: the result is discretely divergence-free,
    G^T M (y + G psi) = 0 on the free H1 dofs, and equals the projection formed from the oracle's assembled matrices."""
    from palace_b200 import capi

    capi.set_stream(b2p_ctx)
    mesh = ts.box_tet_mesh((2, 2, 1), (1.0, 0.8, 0.9), jitter=0.2, scramble_seed=8)
    p = 2
    nd = ts.build_nd_tet_space(mesh, p)
    h1l = [ts.build_h1_tet_space(mesh, nd, q) for q in (1, 2)]
    interp, curl, qpts, qw = ts.nd_tet_tables(p)
    qd = ts.geom_qdata(mesh.node_coords(1), mesh.attr, 1, qpts, qw)
    geom = capi.Geom.general(b2p_ctx, qd)
    eps = cf.coeff_ctx(a=1.7)
    Mop = capi.Op.create_dense(b2p_ctx, geom, O.ND_MASS, nd.ndofs, nd.idx, None, interp, None, eps, curl_orient=nd.curl_orient)
    Mpar = capi.Operator.par(b2p_ctx, nd.ndofs, nd.ndofs, [Mop], None, None, diag_policy=1)   # no essential dofs (b2p.h)
    git = capi.Interp.dense(b2p_ctx, ts.tet_discrete_gradient(p), h1l[1].idx, h1l[1].ndofs, nd.idx, nd.ndofs, out_curl_orient=ts.dual_orient(nd))
    G = capi.Operator.interp(b2p_ctx, git)
    h1_ops, keep = [], []
    for h1 in h1l:
        _, grad = ts.h1_tet_element(h1.p).tabulate(qpts)
        D = capi.Op.create_dense(b2p_ctx, geom, O.H1_DIFFUSION, h1.ndofs, h1.idx, None, None, grad, eps)
        keep.append(D)
        h1_ops.append(capi.Operator.par(b2p_ctx, h1.ndofs, h1.ndofs, [D], None, h1.ess_dofs, diag_policy=1))
    pit = capi.Interp.dense(b2p_ctx, ts.h1_tet_prolongation(1, 2), h1l[0].idx, h1l[0].ndofs, h1l[1].idx, h1l[1].ndofs)
    h1_P = [capi.Operator.interp(b2p_ctx, pit)]
    df = capi.DivFree(b2p_ctx, Mpar, G, h1_ops, h1_P, h1l[1].ess_dofs, 2, tol=1e-12, max_it=200, coarse_type=0)
    y0 = np.random.default_rng(4).standard_normal(nd.ndofs)
    y = _dev(y0)
    df.mult(y)
    assert df.stats()["converged"]
    # oracle-side matrices
    Me = O.element_matrices(O.ND_MASS, interp, curl, None, qd, eps, nd.P)
    M = np.zeros((nd.ndofs, nd.ndofs))
    for e in range(mesh.ne):
        T = nd.dense_T(e)
        M[np.ix_(nd.idx[e], nd.idx[e])] += T.T @ Me[e] @ T
    dual = ts.dual_orient(nd)

    def dualT(e):
        P_ = nd.P
        T = np.zeros((P_, P_))
        for r in range(P_):
            T[r, r] = dual[e, r, 1]
            if r > 0:
                T[r, r - 1] = dual[e, r, 0]
            if r < P_ - 1:
                T[r, r + 1] = dual[e, r, 2]
        return T

    Gm = np.asarray(ts.global_interp_matrix(ts.tet_discrete_gradient(p), h1l[1].idx, None, nd.idx, dualT, h1l[1].ndofs, nd.ndofs).todense())
    free = np.setdiff1d(np.arange(h1l[1].ndofs), h1l[1].ess_dofs)
    S_ = Gm.T @ M @ Gm
    psi = np.zeros(h1l[1].ndofs)
    psi[free] = np.linalg.solve(S_[np.ix_(free, free)], -(Gm.T @ (M @ y0))[free])
    y_ref = y0 + Gm @ psi
    ys = y.cpu().numpy()
    assert np.linalg.norm(ys - y_ref) < 1e-9 * np.linalg.norm(y_ref)
    assert np.abs((Gm.T @ (M @ ys))[free]).max() < 1e-9 * np.abs(Gm.T @ (M @ y0)).max()
