"""GPU: create / use / destroy of every handle type of the C ABI, in the order a host would tear a solver stack down (solvers
before the operators they point at, operators before their local operators, everything before the context). The Python binding
used by the other tests never destroys operator and solver handles, so the destructors (and the ownership rules of include/b2p.h:
multigrid takes its coarse solver, the assembled-matrix wrapper takes its Krylov solver, shared prolongation matrices are
reference counted) are exercised only here; under the AddressSanitizer build of the emulation (make -C tests/emu asan) a double
free or a use after free in these paths is a report."""
import ctypes as C

import numpy as np
import pytest
import scipy.sparse as sparse

from oracle import pyoracle as O
from palace_b200.host import assemble as asm
from palace_b200.host import hexspace as hs
from tests import common

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def _dev(a):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).cuda()


def test_every_handle_type_is_created_used_and_destroyed():
    """Runs in a child process: a crash in a destructor must be a failed test, not the end of the pytest run."""
    import os
    import subprocess
    import sys

    code = ("import os, sys; sys.path.insert(0, %r)\n"
            "if os.environ.get('B2P_EMU_TESTS') == '1':\n"
            "    from tests.emu import emu_mode; emu_mode.enable()\n"
            "from palace_b200 import capi\n"
            "from tests import test_zzz_lifecycle_gpu as t\n"
            "ctx = capi.Ctx(0); t.lifecycle(ctx); ctx.close(); print('LIFECYCLE OK')\n") % common.__file__.rsplit("/tests/", 1)[0]
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=900, env=dict(os.environ))
    assert r.returncode == 0 and "LIFECYCLE OK" in r.stdout, (r.returncode, r.stdout[-400:], r.stderr[-1200:])


def lifecycle(b2p_ctx):
    from palace_b200 import capi

    L = capi.lib()
    for name in ("b2p_operator_destroy", "b2p_solver_destroy", "b2p_coperator_destroy", "b2p_csolver_destroy", "b2p_interp_destroy",
                 "b2p_spmat_destroy", "b2p_csr_destroy"):
        getattr(L, name).restype = None
        getattr(L, name).argtypes = [C.c_void_p]
    ctx = b2p_ctx
    prob = common.make_problem(n=(2, 2, 2), p=2, n_attr=2)
    geom = common.gpu_geom(ctx, prob)
    nd = {1: hs.build_nd_space(prob.mesh, prob.topo, 1), 2: prob.nd}
    h1 = {1: hs.build_h1_space(prob.mesh, prob.topo, 1), 2: prob.h1}
    bk, bm = common.coefficient(O.CURLCURL, 2, "matrix"), common.coefficient(O.ND_MASS, 2, "matrix")
    bh = common.coefficient(O.H1_DIFFUSION, 2, "matrix")
    K2, M2 = common.gpu_op(ctx, geom, prob, O.CURLCURL, bk), common.gpu_op(ctx, geom, prob, O.ND_MASS, bm)
    D2 = common.gpu_op(ctx, geom, prob, O.H1_DIFFUSION, bh)
    t = hs.tables_1d(1, prob.q1d)
    idx, ori = nd[1].native_restriction()
    K1 = K2.coarsen(1, nd[1].ndofs, idx, ori, nd[1].dof_map, t.Bo, t.Bc, t.Gc)
    M1 = M2.coarsen(1, nd[1].ndofs, idx, ori, nd[1].dof_map, t.Bo, t.Bc, t.Gc)
    D1 = D2.coarsen(1, h1[1].ndofs, h1[1].lex_gid.astype(np.int32), None, None, None, t.Bc, t.Gc)
    A = {p: capi.Operator.par(ctx, nd[p].ndofs, nd[p].ndofs, [K, M], [1.0, 2.0], nd[p].ess_dofs, diag_policy=1)
         for p, (K, M) in ((1, (K1, M1)), (2, (K2, M2)))}
    AG = {p: capi.Operator.par(ctx, h1[p].ndofs, h1[p].ndofs, [D], None, h1[p].ess_dofs, diag_policy=1) for p, D in ((1, D1), (2, D2))}
    G = {p: common.gpu_interp(ctx, h1[p], nd[p], asm.gradient_comps(p)) for p in (1, 2)}
    P = common.gpu_interp(ctx, nd[1], nd[2], asm.nd_prolongation_comps(1, 2))
    n = nd[2].ndofs
    b = np.random.default_rng(0).random(n)
    b[nd[2].ess_dofs] = 0.0
    x = torch.zeros(n, dtype=torch.float64, device="cuda")
    # real stack: FGMRES + multigrid (Hiptmair) with a coarse solver on the assembled matrix
    cg = capi.Solver.krylov(ctx, capi.CG, rel_tol=1e-8, max_it=200)
    cj = capi.Solver.jacobi(ctx)
    coarse = capi.Solver.assembled(ctx, cg, cj)
    mg = capi.Solver.gmg(ctx, coarse, [P], [G[1], G[2]], cycle_it=1, smooth_it=1, cheby_order=4)
    mg.gmg_set_operators([A[1], A[2]], [AG[1], AG[2]])
    ks = capi.Solver.krylov(ctx, capi.FGMRES, rel_tol=1e-8, max_it=40)
    ks.set_operator(A[2])
    ks.set_preconditioner(mg)
    ks.mult(_dev(b), x)
    assert ks.stats()["converged"]
    # complex stack on the same local operators
    Z = capi.ComplexOperator.par(ctx, n, n, [K2, M2], [1.0, -0.5 + 0.1j], nd[2].ess_dofs, diag_policy=1)
    W = capi.ComplexOperator.wrap(ctx, A[2], None)
    zj = capi.ComplexSolver.jacobi(ctx)
    zj.set_operator(Z)
    zk = capi.ComplexSolver.krylov(ctx, capi.GMRES, rel_tol=1e-6, max_it=30, max_dim=30)
    zk.set_operator(Z)
    zk.set_preconditioner(zj)
    rp = capi.ComplexSolver.real_pc(ctx, mg)
    xr, xi = torch.zeros_like(x), torch.zeros_like(x)
    zk.mult(_dev(b), _dev(0 * b), xr, xi)
    # general-prolongation operators sharing one sparse matrix
    Pm = capi.SpMat(ctx, sparse.identity(n, format="csr"))
    Kl, Ml = common.gpu_op(ctx, geom, prob, O.CURLCURL, bk), common.gpu_op(ctx, geom, prob, O.ND_MASS, bm)  # (one op = one mask)
    Aloc = capi.Operator.par(ctx, n, n, [Kl, Ml], [1.0, 2.0], None, diag_policy=1)
    R1 = capi.operator_rap(ctx, Aloc, Pm, nd[2].ess_dofs)
    T1 = capi.operator_triple(ctx, Pm, Aloc, Pm)
    y = torch.empty_like(x)
    R1.mult(x, y)
    T1.mult(x, y)
    # ---- tear down: users first ----
    for s_ in (zk, zj, rp):
        L.b2p_csolver_destroy(s_.h)
    for a_ in (Z, W):
        L.b2p_coperator_destroy(a_.h)
    for s_ in (ks, mg, coarse, cg, cj):   # mg owns the coarse solver, the wrapper owns cg / cj: their handles are empty shells now
        L.b2p_solver_destroy(s_.h)
    for o_ in (R1, T1):
        L.b2p_operator_destroy(o_.h)
    L.b2p_spmat_destroy(Pm.h)             # the two operators held references; this drops the caller's
    Pm.h = None
    for o_ in (Aloc, P, G[1], G[2], A[1], A[2], AG[1], AG[2]):
        L.b2p_operator_destroy(o_.h)
    for o_ in (P, G[1], G[2]):
        L.b2p_interp_destroy(o_._keep[0].h)
    # the context is still usable afterwards
    K3 = common.gpu_op(ctx, geom, prob, O.CURLCURL, bk)
    K3.apply(x, y)
    assert bool(torch.isfinite(y).all())
