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
