"""GPU parity of the device full assembly (b2p_csr_*): the coarse-level CSR matrix the reference builds with
CeedOperatorFullAssemble (/root/reference/palace/fem/libceed/operator.cpp:262-523) for HYPRE / sparse direct
solvers, against the oracle's assembled sparse matrices (entry-wise, 1e-12 of the largest entry), its action against the
matrix-free apply, and EliminateBC semantics (rap.cpp:141-146)."""
import numpy as np
import pytest

from oracle import pyoracle as O
from palace_b200.host import coeff as cf
from tests import common

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def _dev(a):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).cuda()


def _maxdiff(A, B):
    D = (A - B).tocoo()
    return (np.abs(D.data).max() if D.nnz else 0.0) / np.abs(B.tocoo().data).max()


@pytest.mark.parametrize("p", [1, 2])
def test_hex_sum_operator_assembles_to_the_oracle_matrix(b2p_ctx, p):
    from palace_b200 import capi

    prob = common.make_problem(n=(3, 2, 2), p=p, n_attr=3)
    geom = common.gpu_geom(b2p_ctx, prob)
    kb = common.coefficient(O.CURLCURL, 3, "matrix")
    mb = common.coefficient(O.ND_MASS, 3, "matrix", a_mass=1.7)
    K, M = common.gpu_op(b2p_ctx, geom, prob, O.CURLCURL, kb), common.gpu_op(b2p_ctx, geom, prob, O.ND_MASS, mb)
    A = capi.Csr(b2p_ctx, K)
    A.assemble([K, M], [1.0, -0.6])
    ref = common.oracle_matrix(prob, O.CURLCURL, kb, eliminate=False) - 0.6 * common.oracle_matrix(prob, O.ND_MASS, mb, eliminate=False)
    got = A.to_scipy()
    assert got.shape == ref.shape and _maxdiff(got, ref.tocsr()) < 1e-12
    assert (np.diff(got.indptr) > 0).all() and all((np.diff(got.indices[a:b]) > 0).all() for a, b in zip(got.indptr[:-1], got.indptr[1:]))
    # the assembled matrix acts like the matrix-free operators
    x = np.random.default_rng(0).random(prob.nd.ndofs)
    y = torch.empty(prob.nd.ndofs, dtype=torch.float64, device="cuda")
    A.mult(_dev(x), y)
    yk, ym = torch.empty_like(y), torch.empty_like(y)
    K.apply(_dev(x), yk)
    M.apply(_dev(x), ym)
    y_mf = yk.cpu().numpy() - 0.6 * ym.cpu().numpy()
    assert np.linalg.norm(y.cpu().numpy() - y_mf) < 1e-12 * np.linalg.norm(y_mf)
    # re-assembly with other coefficients reuses the pattern
    A.assemble([M], [2.0])
    assert _maxdiff(A.to_scipy(), (2.0 * common.oracle_matrix(prob, O.ND_MASS, mb, eliminate=False)).tocsr()) < 1e-12
    # essential rows / columns (DIAG_ONE)
    A.eliminate(prob.nd.ess_dofs, 1)
    want = common.oracle_matrix(prob, O.ND_MASS, mb, eliminate=False).tolil() * 2.0
    ess = prob.nd.ess_dofs
    want[ess, :] = 0
    want[:, ess] = 0
    want[ess, ess] = 1.0
    assert _maxdiff(A.to_scipy(), want.tocsr()) < 1e-12


def test_h1_coarse_matrix(b2p_ctx):
    from palace_b200 import capi

    prob = common.make_problem(n=(3, 3, 2), p=1, n_attr=2)
    geom = common.gpu_geom(b2p_ctx, prob)
    blob = common.coefficient(O.H1_DIFFUSION, 2, "matrix")
    op = common.gpu_op(b2p_ctx, geom, prob, O.H1_DIFFUSION, blob)
    A = capi.Csr(b2p_ctx, op)
    A.assemble([op])
    ref = common.oracle_matrix(prob, O.H1_DIFFUSION, blob, eliminate=False)
    assert _maxdiff(A.to_scipy(), ref.tocsr()) < 1e-12


@pytest.mark.parametrize("p", [1, 2])
def test_tetrahedron_operator_with_curl_oriented_restriction(b2p_ctx, p):
    from palace_b200 import capi
    from palace_b200.host import tetspace as ts

    mesh = ts.box_tet_mesh((2, 1, 1), (1.0, 0.8, 0.9), jitter=0.25, scramble_seed=5, n_attr=1)
    sp = ts.build_nd_tet_space(mesh, p)
    interp, curl, qpts, qw = ts.nd_tet_tables(p)
    qd = ts.geom_qdata(mesh.node_coords(1), mesh.attr, 1, qpts, qw)
    blob = cf.coeff_ctx_pair(cf.coeff_ctx(a=1.0), cf.coeff_ctx(a=2.0))
    geom = capi.Geom.general(b2p_ctx, qd)
    op = capi.Op.create_dense(b2p_ctx, geom, O.CURLCURL_MASS, sp.ndofs, sp.idx, None, interp, curl, blob, curl_orient=sp.curl_orient)
    A = capi.Csr(b2p_ctx, op)
    A.assemble([op])
    got = A.to_scipy().toarray()
    Ae = O.element_matrices(O.CURLCURL_MASS, interp, curl, None, qd, blob, sp.P)
    ref = np.zeros((sp.ndofs, sp.ndofs))
    for e in range(mesh.ne):
        T = sp.dense_T(e)
        ref[np.ix_(sp.idx[e], sp.idx[e])] += T.T @ Ae[e] @ T
    assert np.abs(got - ref).max() < 1e-12 * np.abs(ref).max()


def test_assembled_matrix_as_the_coarse_level_operator(b2p_ctx):
    """The assembled, eliminated coarse matrix behind the operator interface (b2p_operator_csr): same action and diagonal
    as the matrix-free ParOperator, and a Jacobi-PCG on it reproduces the sparse direct solution of the oracle matrix —
    the reference's coarse-level flow (assembled matrix handed to the coarse solver)."""
    import scipy.sparse.linalg as spla

    from palace_b200 import capi

    prob = common.make_problem(n=(3, 3, 2), p=1, n_attr=2)
    geom = common.gpu_geom(b2p_ctx, prob)
    kb = common.coefficient(O.CURLCURL, 2, "matrix")
    mb = common.coefficient(O.ND_MASS, 2, "matrix", a_mass=1.3)
    K, M = common.gpu_op(b2p_ctx, geom, prob, O.CURLCURL, kb), common.gpu_op(b2p_ctx, geom, prob, O.ND_MASS, mb)
    n, ess = prob.nd.ndofs, prob.nd.ess_dofs
    csr = capi.Csr(b2p_ctx, K)
    csr.assemble([K, M], [1.0, 0.8])
    csr.eliminate(ess, 1)
    Ac = capi.Operator.from_csr(b2p_ctx, csr)
    Amf = capi.Operator.par(b2p_ctx, n, n, [K, M], [1.0, 0.8], ess_tdofs=ess, diag_policy=1)
    assert Ac.height == n

    rng = np.random.default_rng(4)
    x = rng.random(n)
    y1, y2 = torch.empty(n, dtype=torch.float64, device="cuda"), torch.empty(n, dtype=torch.float64, device="cuda")
    Ac.mult(_dev(x), y1)
    Amf.mult(_dev(x), y2)
    assert np.linalg.norm(y1.cpu().numpy() - y2.cpu().numpy()) < 1e-12 * np.linalg.norm(y2.cpu().numpy())
    d1, d2 = torch.empty_like(y1), torch.empty_like(y1)
    Ac.assemble_diagonal(d1)
    Amf.assemble_diagonal(d2)
    assert np.abs(d1.cpu().numpy() - d2.cpu().numpy()).max() < 1e-12 * np.abs(d2.cpu().numpy()).max()

    b = rng.random(n)
    b[ess] = 0.0
    J = capi.Solver.jacobi(b2p_ctx)
    J.set_operator(Ac)
    cg = capi.Solver.krylov(b2p_ctx, 0, rel_tol=1e-12, max_it=500)
    cg.set_operator(Ac)
    cg.set_preconditioner(J)
    xd = torch.zeros(n, dtype=torch.float64, device="cuda")
    cg.mult(_dev(b), xd)
    want = spla.spsolve(csr.to_scipy().tocsc(), b)
    assert np.linalg.norm(xd.cpu().numpy() - want) < 1e-9 * np.linalg.norm(want)
