"""GPU: ParOperator with a general conforming prolongation (b2p_operator_rap, b2p_spmat) on a non-conforming hexahedral mesh --
P^T A P with essential true dofs and the |P|^T d_L diagonal of /root/reference/palace/linalg/rap.cpp:154-234 -- against the same
products formed with SciPy from the oracle's assembled local matrix, and a PCG solve of the constrained curl-curl + mass system.
The prolongation comes from the host layer (palace_b200/host/nonconforming.py, checked on the CPU in test_nonconforming_cpu.py)."""
import numpy as np
import pytest
import scipy.sparse.linalg as spla

from oracle import pyoracle as O
from oracle import solvers as S
from palace_b200.host import nonconforming as nc
from tests import common

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu
RTOL = 1e-12


def _dev(a):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).cuda()


def _rel(a, b):
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


@pytest.fixture(scope="module", params=[2, 3])
def setup(request, b2p_ctx):
    from palace_b200 import capi

    p = request.param
    hb = nc.hanging_box_mesh(nc=(1, 2, 1), nfx=2, h=1.0, scramble_seed=7, n_attr=3)
    cs = nc.build_constrained_nd_space(hb, p)
    prob = common.problem_on_mesh(hb.mesh, p)
    kind = O.CURLCURL_MASS
    blob = common.coefficient(kind, 3, "matrix", a_mass=1.3, a_curl=0.7)
    geom = common.gpu_geom(b2p_ctx, prob)
    op = common.gpu_op(b2p_ctx, geom, prob, kind, blob)
    nL = prob.nd.ndofs
    A_loc = capi.Operator.par(b2p_ctx, nL, nL, [op], None, None, diag_policy=1)  # the L-vector operator: no essential dofs
    P = capi.SpMat(b2p_ctx, cs.P)
    A = capi.operator_rap(b2p_ctx, A_loc, P, cs.ess_tdofs, diag_policy=1)
    AL = common.oracle_matrix(prob, kind, blob, eliminate=False)
    return dict(cs=cs, prob=prob, A=A, P=P, AL=AL, A_loc=A_loc, kind=kind, blob=blob)


def test_prolongation_products(setup, b2p_ctx):
    cs, P = setup["cs"], setup["P"]
    nL, nT = cs.P.shape
    rng = np.random.default_rng(0)
    x, z = rng.random(nT) - 0.5, rng.random(nL) - 0.5
    y = torch.empty(nL, dtype=torch.float64, device="cuda")
    P.mult(_dev(x), y)
    assert _rel(y.cpu().numpy(), cs.P @ x) < 1e-14
    yt = torch.empty(nT, dtype=torch.float64, device="cuda")
    P.mult(_dev(z), yt, transpose=True)
    assert _rel(yt.cpu().numpy(), cs.P.T @ z) < 1e-14


def test_rap_operator_matches_the_eliminated_triple_product(setup):
    cs, A, AL = setup["cs"], setup["A"], setup["AL"]
    nT = cs.P.shape[1]
    At = S.eliminate((cs.P.T @ AL @ cs.P).tocsr(), cs.ess_tdofs)
    x = np.random.default_rng(1).random(nT) - 0.5
    y = torch.full((nT,), 7.0, dtype=torch.float64, device="cuda")
    A.mult(_dev(x), y)
    assert _rel(y.cpu().numpy(), At @ x) < RTOL
    A.mult_transpose(_dev(x), y)
    assert _rel(y.cpu().numpy(), At.T @ x) < RTOL
    y2 = _dev(np.ones(nT))
    A.add_mult(_dev(x), y2, -0.5)
    assert _rel(y2.cpu().numpy(), 1.0 - 0.5 * (At @ x)) < RTOL
    assert A.height == nT


def test_rap_diagonal_is_the_absolute_value_product(setup):
    """|P|^T d_L with essential entries set to one (rap.cpp:162-191) -- not diag(P^T A P): the reference's choice for AMR meshes."""
    cs, A, AL = setup["cs"], setup["A"], setup["AL"]
    nT = cs.P.shape[1]
    d_ref = abs(cs.P).T @ AL.diagonal()
    d_ref[cs.ess_tdofs] = 1.0
    d = torch.empty(nT, dtype=torch.float64, device="cuda")
    A.assemble_diagonal(d)
    assert _rel(d.cpu().numpy(), d_ref) < RTOL


def test_pcg_on_the_constrained_system(setup, b2p_ctx):
    """CG + Jacobi (the |P|^T d_L diagonal) on P^T (K + M) P: the solution is the SciPy solve of the eliminated triple product."""
    from palace_b200 import capi

    cs, A, AL = setup["cs"], setup["A"], setup["AL"]
    nT = cs.P.shape[1]
    At = S.eliminate((cs.P.T @ AL @ cs.P).tocsr(), cs.ess_tdofs)
    b = np.random.default_rng(2).random(nT) - 0.5
    b[cs.ess_tdofs] = 0.0
    x_ref = spla.spsolve(At.tocsc(), b)
    pc = capi.Solver.jacobi(b2p_ctx)
    pc.set_operator(A)
    cg = capi.Solver.krylov(b2p_ctx, capi.CG, rel_tol=1e-12, max_it=2000)
    cg.set_operator(A)
    cg.set_preconditioner(pc)
    x = torch.zeros(nT, dtype=torch.float64, device="cuda")
    cg.mult(_dev(b), x)
    st = cg.stats()
    assert st["converged"]
    assert _rel(x.cpu().numpy(), x_ref) < 1e-8


def test_p_multigrid_on_the_non_conforming_mesh(b2p_ctx):
    """Two p-levels {1, 2} on the mesh with hanging faces: level operators P_l^T A_l P_l (b2p_operator_rap), level prolongation
    R_2 I P_1 (b2p_operator_triple: ParOperator(interpolator, use_R), rap.cpp:320-345), Chebyshev smoothing on the |P|^T d_L
    diagonal; one V-cycle against the NumPy multigrid of oracle/solvers.py on the SciPy triple products, and FGMRES + V-cycle
    to the direct solution."""
    import scipy.sparse as sparse

    from palace_b200 import capi
    from palace_b200.host import assemble as asm
    from palace_b200.host import hexspace as hs

    hb = nc.hanging_box_mesh(nc=(1, 2, 1), nfx=2, h=1.0, scramble_seed=9, n_attr=2)
    prob = common.problem_on_mesh(hb.mesh, 2)
    kind = O.CURLCURL_MASS
    blob = common.coefficient(kind, 2, "matrix", a_mass=1.0, a_curl=0.5)
    geom = common.gpu_geom(b2p_ctx, prob)
    nd = {1: hs.build_nd_space(hb.mesh, prob.topo, 1), 2: prob.nd}
    cs = {p: nc.build_constrained_nd_space(hb, p) for p in (1, 2)}
    op = {2: common.gpu_op(b2p_ctx, geom, prob, kind, blob)}
    t = hs.tables_1d(1, prob.q1d)
    idx, ori = nd[1].native_restriction()
    op[1] = op[2].coarsen(1, nd[1].ndofs, idx, ori, nd[1].dof_map, t.Bo, t.Bc, t.Gc)
    Pm, A, At, dref = {}, {}, {}, {}
    keep = []
    for p in (1, 2):
        assert np.array_equal(cs[p].space.lex_gid, nd[p].lex_gid)
        nL = nd[p].ndofs
        Aloc = capi.Operator.par(b2p_ctx, nL, nL, [op[p]], None, None, diag_policy=1)
        Pm[p] = capi.SpMat(b2p_ctx, cs[p].P)
        A[p] = capi.operator_rap(b2p_ctx, Aloc, Pm[p], cs[p].ess_tdofs, diag_policy=1)
        AL = common.oracle_matrix(prob, kind, blob, space=nd[p], eliminate=False)
        At[p] = S.eliminate((cs[p].P.T @ AL @ cs[p].P).tocsr(), cs[p].ess_tdofs)
        dref[p] = abs(cs[p].P).T @ AL.diagonal()
        dref[p][cs[p].ess_tdofs] = 1.0
        keep.append(Aloc)
    # level prolongation on true dofs
    I_loc = common.gpu_interp(b2p_ctx, nd[1], nd[2], asm.nd_prolongation_comps(1, 2))
    tr = np.nonzero(cs[2].true_of >= 0)[0]
    R2 = sparse.csr_matrix((np.ones(tr.size), (cs[2].true_of[tr], tr)), shape=(cs[2].P.shape[1], nd[2].ndofs))
    Plev = capi.operator_triple(b2p_ctx, capi.SpMat(b2p_ctx, R2), I_loc, Pm[1])
    Por = (R2 @ common.oracle_interp(nd[1], nd[2], hs.nd_prolongation_matrix(1, 2)) @ cs[1].P).tocsr()
    rng = np.random.default_rng(4)
    xc = rng.standard_normal(Por.shape[1])
    yf = torch.empty(Por.shape[0], dtype=torch.float64, device="cuda")
    Plev.mult(_dev(xc), yf)
    assert _rel(yf.cpu().numpy(), Por @ xc) < 1e-13
    rf = rng.standard_normal(Por.shape[0])
    xcd = torch.empty(Por.shape[1], dtype=torch.float64, device="cuda")
    Plev.mult_transpose(_dev(rf), xcd)
    assert _rel(xcd.cpu().numpy(), Por.T @ rf) < 1e-13
    # multigrid: CG + Jacobi to 1e-13 on level 0, Chebyshev (order 4) on level 1
    coarse = capi.Solver.krylov(b2p_ctx, 0, rel_tol=1e-13, max_it=2000)
    cj = capi.Solver.jacobi(b2p_ctx)
    cj.set_operator(A[1])
    coarse.set_preconditioner(cj)
    coarse.set_operator(A[1])
    M = capi.Solver.gmg(b2p_ctx, coarse, [Plev], None, cycle_it=1, smooth_it=1, cheby_order=4, sf_max=1.0, sf_min=0.0, fourth_kind=True)
    M.gmg_set_operators([A[1], A[2]], None)
    c1 = capi.Solver.chebyshev(b2p_ctx, 1, 4)
    c1.set_operator(A[2])
    sm = S.ChebSmoother(At[2], c1.lambda_max(), 4)
    sm.dinv = 1.0 / dref[2]
    lu = spla.splu(At[1].tocsc())
    ref = S.Gmg([At[1], At[2]], [Por], [None, sm], lambda v: lu.solve(v), [cs[1].ess_tdofs, cs[2].ess_tdofs])
    n = At[2].shape[0]
    x = rng.standard_normal(n)
    x[cs[2].ess_tdofs] = 0.0
    yd = torch.empty(n, dtype=torch.float64, device="cuda")
    M.mult(_dev(x), yd)
    assert _rel(yd.cpu().numpy(), ref.mult(x)) < 5e-3  # the lambda_max estimate carries 1e-4
    K = capi.Solver.krylov(b2p_ctx, 2, rel_tol=1e-10, max_it=80, max_dim=80)
    K.set_operator(A[2])
    K.set_preconditioner(M)
    b = rng.standard_normal(n)
    b[cs[2].ess_tdofs] = 0.0
    xd = torch.zeros(n, dtype=torch.float64, device="cuda")
    K.mult(_dev(b), xd)
    st = K.stats()
    assert st["converged"], st  # (57 iterations: Chebyshev alone, without the auxiliary-space correction, smooths ND operators poorly)
    assert _rel(xd.cpu().numpy(), spla.spsolve(At[2].tocsc(), b)) < 1e-8


def test_hiptmair_multigrid_on_the_non_conforming_mesh(b2p_ctx):
    """The reference's production preconditioner on the mesh with hanging faces: p-multigrid {1, 2} with Chebyshev + auxiliary-space
    (Hiptmair) smoothing (distrelaxation.cpp:99-151) -- ND and H1 level operators from b2p_operator_rap, discrete gradients
    R_nd G P_h1 and the level prolongation R I P from b2p_operator_triple -- one V-cycle against oracle/solvers.py, and FGMRES
    converging in a handful of iterations where plain Chebyshev smoothing needed 57."""
    from palace_b200 import capi
    from palace_b200.host import assemble as asm
    from palace_b200.host import hexspace as hs

    hb = nc.hanging_box_mesh(nc=(1, 2, 1), nfx=2, h=1.0, scramble_seed=9, n_attr=2)
    prob = common.problem_on_mesh(hb.mesh, 2)
    kind = O.CURLCURL_MASS
    blob = common.coefficient(kind, 2, "matrix", a_mass=1.0, a_curl=0.5)
    blob_h1 = common.coefficient(O.H1_DIFFUSION, 2, "matrix", a_mass=1.0)
    geom = common.gpu_geom(b2p_ctx, prob)
    orders = (1, 2)
    nd = {1: hs.build_nd_space(hb.mesh, prob.topo, 1), 2: prob.nd}
    h1 = {1: hs.build_h1_space(hb.mesh, prob.topo, 1), 2: prob.h1}
    cn = {p: nc.build_constrained_nd_space(hb, p) for p in orders}
    ch = {p: nc.build_constrained_h1_space(hb, p) for p in orders}
    op = {2: common.gpu_op(b2p_ctx, geom, prob, kind, blob)}
    oph = {2: common.gpu_op(b2p_ctx, geom, prob, O.H1_DIFFUSION, blob_h1)}
    t = hs.tables_1d(1, prob.q1d)
    idx, ori = nd[1].native_restriction()
    op[1] = op[2].coarsen(1, nd[1].ndofs, idx, ori, nd[1].dof_map, t.Bo, t.Bc, t.Gc)
    oph[1] = oph[2].coarsen(1, h1[1].ndofs, h1[1].lex_gid.astype(np.int32), None, None, None, t.Bc, t.Gc)
    A, AG, G, At, AGt, Gt, dA, dG, keep = {}, {}, {}, {}, {}, {}, {}, {}, []

    def rap(local_op, cs, nL, AL):
        Aloc = capi.Operator.par(b2p_ctx, nL, nL, [local_op], None, None, diag_policy=1)
        Pm = capi.SpMat(b2p_ctx, cs.P)
        keep.extend([Aloc, Pm])
        d = abs(cs.P).T @ AL.diagonal()
        d[cs.ess_tdofs] = 1.0
        return capi.operator_rap(b2p_ctx, Aloc, Pm, cs.ess_tdofs, diag_policy=1), Pm, S.eliminate((cs.P.T @ AL @ cs.P).tocsr(), cs.ess_tdofs), d

    Pnd, Ph1 = {}, {}
    for p in orders:
        A[p], Pnd[p], At[p], dA[p] = rap(op[p], cn[p], nd[p].ndofs, common.oracle_matrix(prob, kind, blob, space=nd[p], eliminate=False))
        AG[p], Ph1[p], AGt[p], dG[p] = rap(oph[p], ch[p], h1[p].ndofs, common.oracle_matrix(prob, O.H1_DIFFUSION, blob_h1, space=h1[p], eliminate=False))
        Gloc = common.gpu_interp(b2p_ctx, h1[p], nd[p], asm.gradient_comps(p))
        Rn = nc.restriction_matrix(cn[p])
        G[p] = capi.operator_triple(b2p_ctx, capi.SpMat(b2p_ctx, Rn), Gloc, Ph1[p])
        Gt[p] = (Rn @ common.oracle_interp(h1[p], nd[p], hs.discrete_gradient_matrix(p)) @ ch[p].P).tocsr()
        keep.append(Gloc)
    I_loc = common.gpu_interp(b2p_ctx, nd[1], nd[2], asm.nd_prolongation_comps(1, 2))
    R2 = nc.restriction_matrix(cn[2])
    Plev = capi.operator_triple(b2p_ctx, capi.SpMat(b2p_ctx, R2), I_loc, Pnd[1])
    Por = (R2 @ common.oracle_interp(nd[1], nd[2], hs.nd_prolongation_matrix(1, 2)) @ cn[1].P).tocsr()
    rng = np.random.default_rng(5)
    xg = rng.standard_normal(Gt[2].shape[1])
    yg = torch.empty(Gt[2].shape[0], dtype=torch.float64, device="cuda")
    G[2].mult(_dev(xg), yg)
    assert _rel(yg.cpu().numpy(), Gt[2] @ xg) < 1e-13
    order = 4
    coarse = capi.Solver.krylov(b2p_ctx, 0, rel_tol=1e-13, max_it=2000)
    cj = capi.Solver.jacobi(b2p_ctx)
    cj.set_operator(A[1])
    coarse.set_preconditioner(cj)
    coarse.set_operator(A[1])
    M = capi.Solver.gmg(b2p_ctx, coarse, [Plev], [G[1], G[2]], cycle_it=1, smooth_it=1, cheby_order=order, sf_max=1.0, sf_min=0.0,
                        fourth_kind=True)
    M.gmg_set_operators([A[1], A[2]], [AG[1], AG[2]])
    c1, c2 = capi.Solver.chebyshev(b2p_ctx, 1, order), capi.Solver.chebyshev(b2p_ctx, 1, order)
    c1.set_operator(A[2])
    c2.set_operator(AG[2])
    sm = S.DistRelax(At[2], AGt[2], Gt[2], ch[2].ess_tdofs, c1.lambda_max(), c2.lambda_max(), order)
    sm.dinv, sm.dinv_G = 1.0 / dA[2], 1.0 / dG[2]
    lu = spla.splu(At[1].tocsc())
    ref = S.Gmg([At[1], At[2]], [Por], [None, sm], lambda v: lu.solve(v), [cn[1].ess_tdofs, cn[2].ess_tdofs])
    n = At[2].shape[0]
    x = rng.standard_normal(n)
    x[cn[2].ess_tdofs] = 0.0
    yd = torch.empty(n, dtype=torch.float64, device="cuda")
    M.mult(_dev(x), yd)
    assert _rel(yd.cpu().numpy(), ref.mult(x)) < 5e-3
    K = capi.Solver.krylov(b2p_ctx, 2, rel_tol=1e-10, max_it=60, max_dim=60)
    K.set_operator(A[2])
    K.set_preconditioner(M)
    b = rng.standard_normal(n)
    b[cn[2].ess_tdofs] = 0.0
    xd = torch.zeros(n, dtype=torch.float64, device="cuda")
    K.mult(_dev(b), xd)
    st = K.stats()
    assert st["converged"] and st["its"] <= 25, st
    assert _rel(xd.cpu().numpy(), spla.spsolve(At[2].tocsc(), b)) < 1e-8


def test_complex_system_on_the_non_conforming_mesh(b2p_ctx):
    """ComplexWrapperOperator over two general-prolongation operators (real part P^T (K - w^2 M) P, imaginary part P^T (w C) P with
    DIAG_ZERO): complex GMRES with the complex Jacobi smoother solves the lossy constrained system to the SciPy solution; the wrapper
    reports the essential true dofs of its parts (what the complex multigrid masks restricted residuals with)."""
    from palace_b200 import capi

    p = 2
    hb = nc.hanging_box_mesh(nc=(1, 1, 1), nfx=2, h=1.0, scramble_seed=2, n_attr=1)
    cs = nc.build_constrained_nd_space(hb, p)
    prob = common.problem_on_mesh(hb.mesh, p)
    geom = common.gpu_geom(b2p_ctx, prob)
    bk, bm = common.coefficient(O.CURLCURL, 1, "const"), common.coefficient(O.ND_MASS, 1, "const")
    nL = prob.nd.ndofs
    w2, wc = 1.3, 0.4
    ops = [common.gpu_op(b2p_ctx, geom, prob, k, b) for k, b in ((O.CURLCURL, bk), (O.ND_MASS, bm), (O.ND_MASS, bm))]
    Ar_loc = capi.Operator.par(b2p_ctx, nL, nL, ops[:2], [1.0, w2], None, diag_policy=1)
    Ai_loc = capi.Operator.par(b2p_ctx, nL, nL, ops[2:], [wc], None, diag_policy=1)
    P = capi.SpMat(b2p_ctx, cs.P)
    Ar = capi.operator_rap(b2p_ctx, Ar_loc, P, cs.ess_tdofs, diag_policy=1)
    Ai = capi.operator_rap(b2p_ctx, Ai_loc, P, cs.ess_tdofs, diag_policy=0)
    Z = capi.ComplexOperator.wrap(b2p_ctx, Ar, Ai)
    Ko = common.oracle_matrix(prob, O.CURLCURL, bk, eliminate=False)
    Mo = common.oracle_matrix(prob, O.ND_MASS, bm, eliminate=False)
    Zo = (S.eliminate((cs.P.T @ (Ko + w2 * Mo) @ cs.P).tocsr(), cs.ess_tdofs)
          + 1j * S.eliminate((cs.P.T @ (wc * Mo) @ cs.P).tocsr(), cs.ess_tdofs, diag_one=False)).tocsc()
    n = cs.P.shape[1]
    rng = np.random.default_rng(6)
    b = rng.standard_normal(n) + 1j * rng.standard_normal(n)
    b[cs.ess_tdofs] = 0.0
    zj = capi.ComplexSolver.jacobi(b2p_ctx)
    zj.set_operator(Z)
    k = capi.ComplexSolver.krylov(b2p_ctx, capi.GMRES, rel_tol=1e-11, max_it=400, max_dim=400)
    k.set_operator(Z)
    k.set_preconditioner(zj)
    xr, xi = torch.zeros(n, dtype=torch.float64, device="cuda"), torch.zeros(n, dtype=torch.float64, device="cuda")
    k.mult(_dev(b.real), _dev(b.imag), xr, xi)
    assert k.stats()["converged"]
    assert _rel(xr.cpu().numpy() + 1j * xi.cpu().numpy(), spla.spsolve(Zo, b)) < 1e-8
