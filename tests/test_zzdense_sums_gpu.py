"""GPU: sums of dense-basis operators (tets, prisms, boundary faces) collapse the way sums of sum-factorised ones do
(BuildParSumOperator, /root/reference/palace/linalg/rap.cpp:764-829): real sums a0 K + a2 M as ONE dense operator (stacked value +
curl table, summed per-element tensors), complex sums as TWO (real and imaginary coefficient sums; 4 dense applies per complex
matvec instead of 2-4 per term). Hex tables stand in for the element (the kernel is element-agnostic)."""
import numpy as np
import pytest

from oracle import pyoracle as O
from oracle import solvers as S
from tests import common

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu
RTOL = 1e-12


def _dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def _rel(a, b):
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


def _terms(ctx, prob, specs):
    """Dense operators of the given (kind, blob) list on ONE geometry handle (what lets them fuse)."""
    from palace_b200 import capi

    geom = capi.Geom.general(ctx, prob.qdata_ref)
    sp = prob.nd
    interp, curl, _ = O.nd_hex_tables(sp.p, prob.q1d)
    idx, ori = sp.native_restriction()
    ops = []
    for kind, blob in specs:
        ops.append(capi.Op.create_dense(ctx, geom, kind, sp.ndofs, idx, ori, interp if kind != O.CURLCURL else None,
                                        curl if kind != O.ND_MASS else None, blob))
    return ops, sp


def test_dense_terms_fuse_into_one_operator(b2p_ctx):
    """a0 K + a2 M over dense-basis terms (tets, prisms: BuildParSumOperator, rap.cpp:764-829) runs as ONE dense operator with the
    stacked value + curl table and the summed per-element tensors; new coefficients without a rebuild."""
    from palace_b200 import capi

    prob = common.make_problem(n=(3, 3, 2), p=2, n_attr=3)
    bk = common.coefficient(O.CURLCURL, 3, "matrix", a_curl=0.7)
    bm = common.coefficient(O.ND_MASS, 3, "matrix", a_mass=1.3)
    (K, M), sp = _terms(b2p_ctx, prob, [(O.CURLCURL, bk), (O.ND_MASS, bm)])
    A = capi.Operator.par(b2p_ctx, sp.ndofs, sp.ndofs, [K, M], [1.0, -2.5], sp.ess_dofs, diag_policy=1)
    assert A.is_fused()
    Ko = common.oracle_matrix(prob, O.CURLCURL, bk, eliminate=False)
    Mo = common.oracle_matrix(prob, O.ND_MASS, bm, eliminate=False)
    x = np.random.default_rng(5).random(sp.ndofs) - 0.5
    y = torch.empty(sp.ndofs, dtype=torch.float64, device="cuda")
    for c in ([1.0, -2.5], [0.3, 4.0]):
        A.set_coefficients(c)
        A.mult(_dev(x), y)
        ref = S.eliminate((c[0] * Ko + c[1] * Mo).tocsr(), sp.ess_dofs) @ x
        assert _rel(y.cpu().numpy(), ref) < RTOL
    d = torch.empty(sp.ndofs, dtype=torch.float64, device="cuda")
    A.assemble_diagonal(d)
    assert _rel(d.cpu().numpy(), S.eliminate((0.3 * Ko + 4.0 * Mo).tocsr(), sp.ess_dofs).diagonal()) < RTOL


def test_dense_complex_sum_runs_as_two_dense_operators(b2p_ctx):
    """K + i w C - w^2 (1 - i tan d) M over dense terms: the two coefficient sums; Mult, MultHermitianTranspose, new coefficients."""
    from palace_b200 import capi

    prob = common.make_problem(n=(3, 2, 2), p=2, n_attr=2)
    bk = common.coefficient(O.CURLCURL, 2, "matrix", a_curl=0.7)
    bm = common.coefficient(O.ND_MASS, 2, "matrix", a_mass=1.3)
    bc = common.coefficient(O.ND_MASS, 2, "scalar", a_mass=0.4)
    kinds = [O.CURLCURL, O.ND_MASS, O.ND_MASS]
    blobs = [bk, bm, bc]
    ops, sp = _terms(b2p_ctx, prob, list(zip(kinds, blobs)))
    n = sp.ndofs
    mats = [common.oracle_matrix(prob, k, b, eliminate=False) for k, b in zip(kinds, blobs)]
    coefs = [1.0 + 0.0j, -2.3 + 0.2j, 0.0 + 0.9j]
    A = capi.ComplexOperator.par(b2p_ctx, n, n, ops, coefs, sp.ess_dofs, diag_policy=1)
    rng = np.random.default_rng(7)
    x = rng.random(n) - 0.5 + 1j * (rng.random(n) - 0.5)
    yr, yi = torch.empty(n, dtype=torch.float64, device="cuda"), torch.empty(n, dtype=torch.float64, device="cuda")

    def ref(cs, herm=False):
        Z = sum((np.conj(c) if herm else c) * M for c, M in zip(cs, mats)).tolil()
        Z[sp.ess_dofs, :] = 0
        Z[:, sp.ess_dofs] = 0
        Z[sp.ess_dofs, sp.ess_dofs] = 1.0
        return Z.tocsr() @ x

    A.mult(_dev(x.real), _dev(x.imag), yr, yi)
    assert _rel(yr.cpu().numpy() + 1j * yi.cpu().numpy(), ref(coefs)) < RTOL
    assert A.fused_applies() == 1
    A.mult_hermitian_transpose(_dev(x.real), _dev(x.imag), yr, yi)
    assert _rel(yr.cpu().numpy() + 1j * yi.cpu().numpy(), ref(coefs, herm=True)) < RTOL
    coefs2 = [0.5 + 0.0j, -7.0 + 0.0j, 0.0 + 0.0j]  # a lossless frequency: the imaginary sum is skipped
    A.set_coefficients(coefs2)
    A.mult(_dev(x.real), _dev(x.imag), yr, yi)
    assert _rel(yr.cpu().numpy() + 1j * yi.cpu().numpy(), ref(coefs2)) < RTOL


def test_fused_dense_sum_assembles_into_the_coarse_matrix(b2p_ctx):
    """The multigrid's assembled coarse level (MfemWrapperSolver analogue) on a ParOperator whose terms were fused into one dense
    operator: the device CSR matrix is the eliminated sum, and the Jacobi-PCG on it solves the system."""
    import scipy.sparse.linalg as spla

    from palace_b200 import capi

    prob = common.make_problem(n=(3, 2, 2), p=1, n_attr=2)
    bk = common.coefficient(O.CURLCURL, 2, "matrix", a_curl=0.7)
    bm = common.coefficient(O.ND_MASS, 2, "matrix", a_mass=1.3)
    (K, M), sp = _terms(b2p_ctx, prob, [(O.CURLCURL, bk), (O.ND_MASS, bm)])
    A = capi.Operator.par(b2p_ctx, sp.ndofs, sp.ndofs, [K, M], [1.0, 2.0], sp.ess_dofs, diag_policy=1)
    assert A.is_fused()
    Ao = S.eliminate((common.oracle_matrix(prob, O.CURLCURL, bk, eliminate=False)
                      + 2.0 * common.oracle_matrix(prob, O.ND_MASS, bm, eliminate=False)).tocsr(), sp.ess_dofs)
    cg = capi.Solver.krylov(b2p_ctx, capi.CG, rel_tol=1e-12, max_it=500)
    solver = capi.Solver.assembled(b2p_ctx, cg, capi.Solver.jacobi(b2p_ctx))
    solver.set_operator(A)
    assert solver.assembled_nnz() >= Ao.nnz
    b = np.random.default_rng(9).random(sp.ndofs)
    b[sp.ess_dofs] = 0.0
    x = torch.zeros(sp.ndofs, dtype=torch.float64, device="cuda")
    solver.mult(_dev(b), x)
    assert _rel(x.cpu().numpy(), spla.spsolve(Ao.tocsc(), b)) < 1e-9


def test_terms_with_different_tables_are_not_fused(b2p_ctx):
    """Two dense operators on one geometry handle that tabulate different bases must stay two operators."""
    from palace_b200 import capi

    prob = common.make_problem(n=(3, 2, 2), p=1, n_attr=1)
    geom = capi.Geom.general(b2p_ctx, prob.qdata_ref)
    sp = prob.nd
    interp, curl, _ = O.nd_hex_tables(sp.p, prob.q1d)
    idx, ori = sp.native_restriction()
    bm = common.coefficient(O.ND_MASS, 1, "const")
    M1 = capi.Op.create_dense(b2p_ctx, geom, O.ND_MASS, sp.ndofs, idx, ori, interp, None, bm)
    M2 = capi.Op.create_dense(b2p_ctx, geom, O.ND_MASS, sp.ndofs, idx, ori, 2.0 * interp, None, bm)
    A = capi.Operator.par(b2p_ctx, sp.ndofs, sp.ndofs, [M1, M2], [1.0, 1.0], None, diag_policy=1)
    assert not A.is_fused()
    x = np.random.default_rng(1).random(sp.ndofs)
    y = torch.empty(sp.ndofs, dtype=torch.float64, device="cuda")
    A.mult(_dev(x), y)
    Mo = common.oracle_matrix(prob, O.ND_MASS, bm, eliminate=False)
    assert _rel(y.cpu().numpy(), 5.0 * (Mo @ x)) < RTOL
