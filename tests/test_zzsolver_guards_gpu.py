"""GPU: edge cases of the Krylov layer found while sweeping a driven problem (tools/driven_sweep_bench.py).
  * a zero right-hand side -- what the imaginary part of a real excitation is when a real preconditioner is applied to both parts
    (PCMatReal, /root/reference/palace/models/spaceoperator.cpp:1098-1105) and the multigrid's coarse solver is a CG -- is a
    converged solve with x = 0, not 0 / 0 in the first step length;
  * FGMRES without a preconditioner is an error (iterative.cpp:738 MFEM_VERIFY), not a null dereference."""
import numpy as np
import pytest

from oracle import pyoracle as O
from tests import common

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def _dev(a):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).cuda()


@pytest.fixture(scope="module")
def system(b2p_ctx):
    prob = common.make_problem(n=(2, 2, 2), p=2, n_attr=1)
    geom = common.gpu_geom(b2p_ctx, prob)
    blob = common.coefficient(O.CURLCURL_MASS, 1, "const")
    A = common.gpu_par_operator(b2p_ctx, geom, prob, O.CURLCURL_MASS, blob)
    return prob, A


@pytest.mark.parametrize("check_every", [1, 4])
def test_cg_with_a_zero_right_hand_side_converges_to_zero(b2p_ctx, system, check_every):
    from palace_b200 import capi

    prob, A = system
    n = prob.nd.ndofs
    pc = capi.Solver.jacobi(b2p_ctx)
    pc.set_operator(A)
    cg = capi.Solver.krylov(b2p_ctx, capi.CG, rel_tol=1e-3, max_it=50)
    if check_every > 1:
        cg.set_check_interval(check_every)
    cg.set_operator(A)
    cg.set_preconditioner(pc)
    x = torch.full((n,), 3.0, dtype=torch.float64, device="cuda")
    cg.mult(_dev(np.zeros(n)), x)
    st = cg.stats()
    assert st["converged"] and st["its"] == 0
    assert float(x.abs().max()) == 0.0


def test_real_preconditioner_on_a_purely_real_complex_vector(b2p_ctx, system):
    """real_pc(CG) applied to (b, 0): the imaginary part stays exactly zero and finite."""
    from palace_b200 import capi

    prob, A = system
    n = prob.nd.ndofs
    pc = capi.Solver.jacobi(b2p_ctx)
    pc.set_operator(A)
    cg = capi.Solver.krylov(b2p_ctx, capi.CG, rel_tol=1e-8, max_it=400)
    cg.set_operator(A)
    cg.set_preconditioner(pc)
    zpc = capi.ComplexSolver.real_pc(b2p_ctx, cg)
    b = np.random.default_rng(0).random(n)
    b[prob.nd.ess_dofs] = 0.0
    yr, yi = torch.empty(n, dtype=torch.float64, device="cuda"), torch.full((n,), 5.0, dtype=torch.float64, device="cuda")
    zpc.mult(_dev(b), _dev(np.zeros(n)), yr, yi)
    assert bool(torch.isfinite(yr).all()) and float(yi.abs().max()) == 0.0


def test_fgmres_without_a_preconditioner_is_an_error(b2p_ctx, system):
    from palace_b200 import capi

    prob, A = system
    n = prob.nd.ndofs
    k = capi.Solver.krylov(b2p_ctx, capi.FGMRES, rel_tol=1e-6, max_it=5)
    k.set_operator(A)
    with pytest.raises(capi.B2PError):
        k.mult(_dev(np.ones(n)), torch.zeros(n, dtype=torch.float64, device="cuda"))


@pytest.mark.parametrize("kind", [1, 2])
def test_gmres_with_a_zero_right_hand_side_converges_to_zero(b2p_ctx, system, kind):
    from palace_b200 import capi

    prob, A = system
    n = prob.nd.ndofs
    pc = capi.Solver.jacobi(b2p_ctx)
    pc.set_operator(A)
    k = capi.Solver.krylov(b2p_ctx, kind, rel_tol=1e-6, max_it=20)
    k.set_operator(A)
    k.set_preconditioner(pc)
    x = torch.full((n,), 3.0, dtype=torch.float64, device="cuda")
    k.mult(_dev(np.zeros(n)), x)
    assert k.stats()["converged"] and float(x.abs().max()) == 0.0
