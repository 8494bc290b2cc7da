"""GPU parity of opt-in solver variants that are not the reference's literal iteration: the CG with device-resident
recurrence scalars (b2p_solver_krylov_set_check_interval: no host synchronisation per iteration, residual read back every k
iterations) must reach the same solution as the reference iteration (iterative.cpp:361-486) and as a sparse direct solve."""
import numpy as np
import pytest
import scipy.sparse.linalg as spla

from oracle import pyoracle as O
from tests import common

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def _dev(a):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).cuda()


@pytest.mark.parametrize("initial_guess", [False, True])
def test_cg_with_device_scalars(b2p_ctx, initial_guess):
    from palace_b200 import capi

    capi.set_stream(b2p_ctx)
    prob = common.make_problem(n=(3, 2, 2), p=2, n_attr=2)
    geom = common.gpu_geom(b2p_ctx, prob)
    blob = common.coefficient(O.CURLCURL_MASS, 2, "const", a_mass=2.0)
    A = common.gpu_par_operator(b2p_ctx, geom, prob, O.CURLCURL_MASS, blob)
    Ao = common.oracle_matrix(prob, O.CURLCURL_MASS, blob, eliminate=True)
    n = prob.nd.ndofs
    rng = np.random.default_rng(0)
    b = rng.random(n)
    x0 = rng.random(n) if initial_guess else np.zeros(n)
    out = {}
    for k in (1, 5):
        cg = capi.Solver.krylov(b2p_ctx, capi.CG, rel_tol=1e-11, max_it=2000)
        J = capi.Solver.jacobi(b2p_ctx)
        J.set_operator(A)
        cg.set_preconditioner(J)
        cg.set_operator(A)
        cg.set_initial_guess(initial_guess)
        cg.set_check_interval(k)
        x = _dev(x0)
        cg.mult(_dev(b), x)
        st = cg.stats()
        assert st["converged"]
        out[k] = (x.cpu().numpy(), st["its"])
    x_ref = spla.spsolve(Ao.tocsc(), b)
    for k in (1, 5):
        assert np.linalg.norm(out[k][0] - x_ref) < 1e-8 * np.linalg.norm(x_ref)
    # at most check_every - 1 iterations past convergence (plus a few either way: the scatter-add order is not fixed, so two
    # runs of a 100+ iteration CG do not follow bit-identical paths)
    assert abs(out[5][1] - out[1][1]) < 20 and out[5][1] % 5 == 0
