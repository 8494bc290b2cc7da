"""GPU parity of the Nedelec TETRAHEDRON path (BASELINE configs 3 and 4 are tet meshes): the dense-basis
DMMA operator (b2p_op_create_dense) fed with real simplex tables, the tridiagonal curl-oriented restriction
of scrambled tets (/root/reference/palace/fem/libceed/restriction.cpp:301-368) and prebuilt q-data of straight
and curved tets, against the oracle. Bit-level agreement is not expected (different summation order):
||dy|| <= 1e-12 ||y||, the reference's own bound (test/unit/test-libceed.cpp:262-268)."""
import numpy as np
import pytest

from oracle import pyoracle as O
from palace_b200.host import coeff as cf
from palace_b200.host import tetspace as ts

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu
RTOL = 1e-12


def _dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def _rel(a, b):
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


def _problem(p, n, geom_order=1, warp=0.0, n_attr=3, degree=None, seed=5):
    mesh = ts.box_tet_mesh(n, (1.0, 0.8, 0.9), jitter=0.25, scramble_seed=seed, n_attr=n_attr, warp_amp=warp)
    sp = ts.build_nd_tet_space(mesh, p)
    interp, curl, qpts, qw = ts.nd_tet_tables(p, degree)
    qd = ts.geom_qdata(mesh.node_coords(geom_order), mesh.attr, geom_order, qpts, qw)
    return mesh, sp, interp, curl, qd


def _blob(kind, n_attr):
    am, mc = cf.test_suite_coefficient(n_attr, "matrix")
    first = cf.coeff_ctx(am, mc, a=1.3)
    second = cf.coeff_ctx(am, mc[::-1].copy() if len(mc) > 1 else mc, a=0.7, transpose=True)
    return cf.coeff_ctx_pair(first, second) if kind == O.CURLCURL_MASS else (second if kind == O.CURLCURL else first)


def _op(ctx, kind, sp, interp, curl, qd, blob, signs_only=False):
    from palace_b200 import capi

    geom = capi.Geom.general(ctx, qd)
    if signs_only:
        return capi.Op.create_dense(ctx, geom, kind, sp.ndofs, sp.idx, sp.orient_signs(), interp, curl, blob)
    return capi.Op.create_dense(ctx, geom, kind, sp.ndofs, sp.idx, None, interp, curl, blob, curl_orient=sp.curl_orient)


@pytest.mark.parametrize("kind", [O.CURLCURL, O.ND_MASS, O.CURLCURL_MASS])
@pytest.mark.parametrize("p", [1, 2, 3])
def test_tet_apply_matches_oracle(b2p_ctx, p, kind):
    mesh, sp, interp, curl, qd = _problem(p, (2, 2, 1))       # 24 tets: three full batches of 8
    blob = _blob(kind, 3)
    op = _op(b2p_ctx, kind, sp, interp, curl, qd, blob)
    x = np.random.default_rng(1).random(sp.ndofs)
    y_ref = O.apply_add_co(kind, interp, curl, sp.idx, sp.curl_orient, qd, blob, x, np.zeros(sp.ndofs))
    yd = torch.full((sp.ndofs,), 7.0, dtype=torch.float64, device="cuda")
    op.apply(_dev(x), yd)                                      # Mult: zero-fill, then add
    assert _rel(yd.cpu().numpy(), y_ref) < RTOL
    op.apply_add_ex(-0.25, _dev(x), yd)                        # AddMult with a coefficient
    assert _rel(yd.cpu().numpy(), 0.75 * y_ref) < RTOL


@pytest.mark.parametrize("p,kind", [(1, O.CURLCURL_MASS), (2, O.CURLCURL), (3, O.ND_MASS), (3, O.CURLCURL_MASS), (6, O.CURLCURL_MASS)])
def test_fused_chunked_dense_kernel_matches_oracle_and_the_round1_kernel(b2p_ctx, p, kind):
    """dense_apply2_kernel (32 elements per CTA, the two GEMMs fused over chunks of 32 points, output tiles in registers)
    takes every batch of >= 32 elements: 54 tets = one full CTA + a ragged one of 22, rules of 4 ... 343 points (ragged
    last chunk), 1 ... 27 dof tiles (one to four per warp), all three kinds (3 or 6 row blocks per chunk)."""
    mesh, sp, interp, curl, qd = _problem(p, (3, 3, 1))       # 54 tets
    blob = _blob(kind, 3)
    op = _op(b2p_ctx, kind, sp, interp, curl, qd, blob)
    x = np.random.default_rng(7).random(sp.ndofs)
    y_ref = O.apply_add_co(kind, interp, curl, sp.idx, sp.curl_orient, qd, blob, x, np.zeros(sp.ndofs))
    yd = torch.full((sp.ndofs,), -3.0, dtype=torch.float64, device="cuda")
    op.apply(_dev(x), yd)
    assert _rel(yd.cpu().numpy(), y_ref) < RTOL
    # the same operator through the element range API in pieces of < 32 elements runs the round-1 kernel
    n_owned = sp.ndofs
    y2 = torch.zeros(sp.ndofs, dtype=torch.float64, device="cuda")
    for e0 in range(0, mesh.ne, 18):
        op.apply_add_split(1.0, _dev(x), None, y2, None, n_owned, e0, min(18, mesh.ne - e0))
    assert _rel(y2.cpu().numpy(), y_ref) < RTOL


def test_tet_p1_sign_orientation_equals_curl_orientation(b2p_ctx):
    mesh, sp, interp, curl, qd = _problem(1, (2, 1, 1))       # 12 tets: ragged last batch
    kind, blob = O.CURLCURL_MASS, _blob(O.CURLCURL_MASS, 3)
    a = _op(b2p_ctx, kind, sp, interp, curl, qd, blob, signs_only=True)
    b = _op(b2p_ctx, kind, sp, interp, curl, qd, blob)
    x = _dev(np.random.default_rng(2).random(sp.ndofs))
    ya, yb = torch.empty_like(x), torch.empty_like(x)
    a.apply(x, ya)
    b.apply(x, yb)
    assert _rel(ya.cpu().numpy(), yb.cpu().numpy()) < RTOL
    y_ref = O.apply_add(kind, interp, curl, sp.idx, sp.orient_signs(), qd, blob, x.cpu().numpy(), np.zeros(sp.ndofs))
    assert _rel(ya.cpu().numpy(), y_ref) < RTOL


def test_curved_tets_p3(b2p_ctx):
    """Quadratic (curved) tets, over-integrated: non-constant Jacobians through the prebuilt q-data path."""
    mesh, sp, interp, curl, qd = _problem(3, (1, 2, 1), geom_order=2, warp=0.03, degree=8)
    kind, blob = O.CURLCURL_MASS, _blob(O.CURLCURL_MASS, 3)
    op = _op(b2p_ctx, kind, sp, interp, curl, qd, blob)
    x = np.random.default_rng(3).random(sp.ndofs)
    y_ref = O.apply_add_co(kind, interp, curl, sp.idx, sp.curl_orient, qd, blob, x, np.zeros(sp.ndofs))
    yd = torch.empty(sp.ndofs, dtype=torch.float64, device="cuda")
    op.apply(_dev(x), yd)
    assert _rel(yd.cpu().numpy(), y_ref) < RTOL


def test_tet_p6_apply_and_symmetry(b2p_ctx):
    """BASELINE config 4's order: P = 216 dofs per tet, 343-point rule (largest element the dense kernel stages)."""
    mesh, sp, interp, curl, qd = _problem(6, (1, 1, 1), n_attr=1)
    kind = O.CURLCURL_MASS
    blob = cf.coeff_ctx_pair(cf.coeff_ctx(a=1.0), cf.coeff_ctx(a=1.0))
    op = _op(b2p_ctx, kind, sp, interp, curl, qd, blob)
    rng = np.random.default_rng(4)
    x, z = rng.random(sp.ndofs), rng.random(sp.ndofs)
    y_ref = O.apply_add_co(kind, interp, curl, sp.idx, sp.curl_orient, qd, blob, x, np.zeros(sp.ndofs))
    yx, yz = torch.empty(sp.ndofs, dtype=torch.float64, device="cuda"), torch.empty(sp.ndofs, dtype=torch.float64, device="cuda")
    op.apply(_dev(x), yx)
    op.apply(_dev(z), yz)
    assert _rel(yx.cpu().numpy(), y_ref) < 1e-11
    assert abs(z @ yx.cpu().numpy() - x @ yz.cpu().numpy()) < 1e-11 * abs(z @ yx.cpu().numpy())


def test_tet_essential_rows_and_pec_box_energy(b2p_ctx):
    """ParOperator over the tet operator with the PEC boundary eliminated (rap.cpp:207-233 semantics): rows of
    essential dofs return x, interior rows ignore essential inputs."""
    from palace_b200 import capi

    mesh, sp, interp, curl, qd = _problem(2, (2, 2, 2), n_attr=1)
    kind = O.CURLCURL_MASS
    blob = cf.coeff_ctx_pair(cf.coeff_ctx(a=1.0), cf.coeff_ctx(a=1.0))
    op = _op(b2p_ctx, kind, sp, interp, curl, qd, blob)
    A = capi.Operator.par(b2p_ctx, sp.ndofs, sp.ndofs, [op], None, sp.ess_dofs, diag_policy=1)
    x = np.random.default_rng(5).random(sp.ndofs)
    xm = x.copy()
    xm[sp.ess_dofs] = 0.0
    y_ref = O.apply_add_co(kind, interp, curl, sp.idx, sp.curl_orient, qd, blob, xm, np.zeros(sp.ndofs))
    y_ref[sp.ess_dofs] = x[sp.ess_dofs]
    yd = torch.empty(sp.ndofs, dtype=torch.float64, device="cuda")
    A.mult(_dev(x), yd)
    assert _rel(yd.cpu().numpy(), y_ref) < RTOL


# ------------------------------------------------------------------------------------------------
# Transfer operators and the multigrid loop on tetrahedra
# ------------------------------------------------------------------------------------------------


def _dual_T(dual, e):
    P = dual.shape[1]
    co = dual[e]
    T = np.zeros((P, P))
    T[np.arange(P), np.arange(P)] = co[:, 1]
    T[np.arange(1, P), np.arange(P - 1)] = co[1:, 0]
    T[np.arange(P - 1), np.arange(1, P)] = co[:-1, 2]
    return T


@pytest.mark.parametrize("p", [2, 3])
def test_tet_interpolators_match_assembled_matrices(b2p_ctx, p):
    """P_l (ND p-1 -> p), G (H1 p -> ND p) and the H1 prolongation as element-dense interpolators with curl-oriented
    restrictions on both sides, Mult and MultTranspose, against the oracle-side assembled matrices."""
    from palace_b200 import capi

    mesh = ts.box_tet_mesh((2, 2, 1), (1.0, 0.8, 0.9), jitter=0.25, scramble_seed=5)
    ndf, ndc = ts.build_nd_tet_space(mesh, p), ts.build_nd_tet_space(mesh, p - 1)
    h1f, h1c = ts.build_h1_tet_space(mesh, ndf, p), ts.build_h1_tet_space(mesh, ndf, p - 1)
    dual = ts.dual_orient(ndf)
    cases = [
        ("P_l", ts.nd_tet_prolongation(p - 1, p), ndc.idx, ndc.ndofs, ndc.curl_orient, ndc.dense_T, ndf.idx, ndf.ndofs, dual),
        ("G", ts.tet_discrete_gradient(p), h1f.idx, h1f.ndofs, None, None, ndf.idx, ndf.ndofs, dual),
        ("P_h1", ts.h1_tet_prolongation(p - 1, p), h1c.idx, h1c.ndofs, None, None, h1f.idx, h1f.ndofs, None),
    ]
    rng = np.random.default_rng(11)
    for name, M, iidx, nin, ico, iT, oidx, nout, oco in cases:
        it = capi.Interp.dense(b2p_ctx, M, iidx, nin, oidx, nout, in_curl_orient=ico, out_curl_orient=oco)
        ref = ts.global_interp_matrix(M, iidx, iT, oidx, (lambda e: _dual_T(oco, e)) if oco is not None else None, nin, nout)
        x, z = rng.random(nin), rng.random(nout)
        y = torch.zeros(nout, dtype=torch.float64, device="cuda")
        it.apply_add(_dev(x), y)
        assert _rel(y.cpu().numpy(), ref @ x) < RTOL, name
        w = torch.zeros(nin, dtype=torch.float64, device="cuda")
        it.apply_add(_dev(z), w, transpose=True, alpha=-2.0)
        assert _rel(w.cpu().numpy(), -2.0 * (ref.T @ z)) < RTOL, name + "^T"


def test_tet_multigrid_preconditioned_solve(b2p_ctx):
    """FGMRES preconditioned by the p-multigrid V-cycle with Hiptmair auxiliary-space smoothing (Chebyshev on ND and on
    the H1 space through the discrete gradient, distrelaxation.cpp:99-151; CG + Jacobi on the coarsest level) on a
    scrambled tetrahedral mesh with PEC boundary: converges, and the answer solves the oracle's assembled system."""
    from palace_b200 import capi

    capi.set_stream(b2p_ctx)
    mesh = ts.box_tet_mesh((2, 1, 2), (1.0, 0.8, 0.9), jitter=0.2, scramble_seed=3)
    orders = [1, 2]
    spaces = [ts.build_nd_tet_space(mesh, p) for p in orders]
    h1s = [ts.build_h1_tet_space(mesh, sp, sp.p) for sp in spaces]
    _, _, qpts, qw = ts.nd_tet_tables(orders[-1])           # every level integrates with the fine rule (shared q-data)
    qd = ts.geom_qdata(mesh.node_coords(1), mesh.attr, 1, qpts, qw)
    geom = capi.Geom.general(b2p_ctx, qd)
    one = cf.coeff_ctx(a=1.0)
    blob = cf.coeff_ctx_pair(one, one)
    pars, auxs, grads = [], [], []
    for sp, h1 in zip(spaces, h1s):
        interp, curl = ts.nd_tet_element(sp.p).tabulate(qpts)
        op = capi.Op.create_dense(b2p_ctx, geom, O.CURLCURL_MASS, sp.ndofs, sp.idx, None, interp, curl, blob, curl_orient=sp.curl_orient)
        pars.append(capi.Operator.par(b2p_ctx, sp.ndofs, sp.ndofs, [op], None, sp.ess_dofs, diag_policy=1))
        # auxiliary operator G^T (K + M) G = H1 diffusion with the mass coefficient
        _, grad = ts.h1_tet_element(h1.p).tabulate(qpts)
        aop = capi.Op.create_dense(b2p_ctx, geom, O.H1_DIFFUSION, h1.ndofs, h1.idx, None, None, grad, one)
        auxs.append(capi.Operator.par(b2p_ctx, h1.ndofs, h1.ndofs, [aop], None, h1.ess_dofs, diag_policy=1))
        git = capi.Interp.dense(b2p_ctx, ts.tet_discrete_gradient(sp.p), h1.idx, h1.ndofs, sp.idx, sp.ndofs, out_curl_orient=ts.dual_orient(sp))
        grads.append(capi.Operator.interp(b2p_ctx, git))
    prol = []
    for c, f in zip(spaces[:-1], spaces[1:]):
        it = capi.Interp.dense(b2p_ctx, ts.nd_tet_prolongation(c.p, f.p), c.idx, c.ndofs, f.idx, f.ndofs,
                               in_curl_orient=c.curl_orient, out_curl_orient=ts.dual_orient(f))
        prol.append(capi.Operator.interp(b2p_ctx, it))
    coarse = capi.Solver.krylov(b2p_ctx, capi.CG, rel_tol=1e-12, max_it=500)
    cj = capi.Solver.jacobi(b2p_ctx)
    cj.set_operator(pars[0])
    coarse.set_preconditioner(cj)
    coarse.set_operator(pars[0])
    gmg = capi.Solver.gmg(b2p_ctx, coarse, prol, grads, cycle_it=1, smooth_it=1, cheby_order=4)
    gmg.gmg_set_operators(pars, auxs)
    ksp = capi.Solver.krylov(b2p_ctx, capi.FGMRES, rel_tol=1e-10, max_it=60)
    ksp.set_preconditioner(gmg)
    ksp.set_operator(pars[-1])
    fine = spaces[-1]
    b = np.random.default_rng(2).random(fine.ndofs)
    b[fine.ess_dofs] = 0.0
    x = torch.zeros(fine.ndofs, dtype=torch.float64, device="cuda")
    ksp.mult(_dev(b), x)
    st = ksp.stats()
    assert st["converged"] and st["its"] < 30, st
    # oracle matrix with the essential rows/columns eliminated (DIAG_ONE)
    interp, curl = ts.nd_tet_element(fine.p).tabulate(qpts)
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
    assert np.linalg.norm(b - A @ xs) < 1e-9 * np.linalg.norm(b)      # true residual against the oracle matrix
    assert _rel(xs, np.linalg.solve(A, b)) < 1e-9 * np.linalg.cond(A)  # error bound: residual x condition number


def test_tet_diagonal_is_exact_with_curl_oriented_restriction(b2p_ctx):
    mesh, sp, interp, curl, qd = _problem(2, (2, 1, 1), n_attr=1)
    blob = cf.coeff_ctx_pair(cf.coeff_ctx(a=1.0), cf.coeff_ctx(a=2.0))
    op = _op(b2p_ctx, O.CURLCURL_MASS, sp, interp, curl, qd, blob)
    d = torch.zeros(sp.ndofs, dtype=torch.float64, device="cuda")
    op.diag_add(d)
    Ae = O.element_matrices(O.CURLCURL_MASS, interp, curl, None, qd, blob, sp.P)
    ref = np.zeros(sp.ndofs)
    for e in range(mesh.ne):
        T = sp.dense_T(e)
        np.add.at(ref, sp.idx[e], np.diag(T.T @ Ae[e] @ T))
    assert _rel(d.cpu().numpy(), ref) < RTOL


def test_dense_operator_coarsening_shares_the_fine_coefficient(b2p_ctx):
    """b2p_op_coarsen_dense (ceed::CeedOperatorCoarsen, fem/libceed/operator.cpp:525-585): the p = 1 level built from the p = 2
    operator acts exactly like a p = 1 operator created with the same geometry and coefficient blob."""
    from palace_b200 import capi

    mesh = ts.box_tet_mesh((2, 1, 2), (1.0, 0.8, 0.9), jitter=0.2, scramble_seed=5, n_attr=2)
    fine_sp, coarse_sp = ts.build_nd_tet_space(mesh, 2), ts.build_nd_tet_space(mesh, 1)
    _, _, qpts, qw = ts.nd_tet_tables(2)
    qd = ts.geom_qdata(mesh.node_coords(1), mesh.attr, 1, qpts, qw)
    geom = capi.Geom.general(b2p_ctx, qd)
    blob = _blob(O.CURLCURL_MASS, 2)
    fi, fc = ts.nd_tet_element(2).tabulate(qpts)
    ci, cc = ts.nd_tet_element(1).tabulate(qpts)
    fine = capi.Op.create_dense(b2p_ctx, geom, O.CURLCURL_MASS, fine_sp.ndofs, fine_sp.idx, None, fi, fc, blob, curl_orient=fine_sp.curl_orient)
    direct = capi.Op.create_dense(b2p_ctx, geom, O.CURLCURL_MASS, coarse_sp.ndofs, coarse_sp.idx, None, ci, cc, blob,
                                  curl_orient=coarse_sp.curl_orient)
    coarse = fine.coarsen_dense(coarse_sp.ndofs, coarse_sp.idx, None, ci, cc, curl_orient=coarse_sp.curl_orient)
    x = np.random.default_rng(2).random(coarse_sp.ndofs)
    y1 = torch.empty(coarse_sp.ndofs, dtype=torch.float64, device="cuda")
    y2 = torch.empty_like(y1)
    coarse.apply(_dev(x), y1)
    direct.apply(_dev(x), y2)
    assert np.linalg.norm(y1.cpu().numpy()) > 0
    assert np.abs(y1.cpu().numpy() - y2.cpu().numpy()).max() <= 1e-13 * np.abs(y2.cpu().numpy()).max()
    d1, d2 = torch.zeros_like(y1), torch.zeros_like(y1)
    coarse.diag_add(d1)
    direct.diag_add(d2)
    assert np.abs(d1.cpu().numpy() - d2.cpu().numpy()).max() <= 1e-13 * np.abs(d2.cpu().numpy()).max()
    fine.close()   # the coarse level keeps the shared coefficient alive
    coarse.apply(_dev(x), y1)
    assert np.abs(y1.cpu().numpy() - y2.cpu().numpy()).max() <= 1e-13 * np.abs(y2.cpu().numpy()).max()
