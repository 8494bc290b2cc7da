"""GPU: PEC box-cavity eigenmodes through the whole stack -- the BASELINE metric's 'eigenfrequency
rel-err' leg in miniature. ARPACK (shift-and-invert, as Palace drives it with host vectors,
/root/reference/palace/linalg/arpack.cpp:631-659) calls the device-resident FGMRES + p-multigrid solve of
(K - sigma M) x = M v; the eigenvalues must agree (a) with the same computation on the oracle's assembled
matrices (sparse LU) to 1e-8 relative -- the north-star tolerance -- and (b) with the closed form
k^2 = pi^2 ((m/a)^2 + (n/b)^2 + (l/d)^2) (docs/src/examples/cylinder.md analogue) to
discretisation accuracy."""
import numpy as np
import pytest
import scipy.sparse.linalg as spla

from oracle import pyoracle as O
from palace_b200.host import assemble as asm
from palace_b200.host import coeff as cf
from palace_b200.host import hexspace as hs
from tests import common

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def _analytic(size, count):
    a, b, d = size
    vals = []
    for m in range(0, 5):
        for n in range(0, 5):
            for l in range(0, 5):
                nz = (m > 0) + (n > 0) + (l > 0)
                if nz < 2:
                    continue
                mult = 2 if nz == 3 else 1  # two polarisations when all indices are non-zero
                vals += [np.pi ** 2 * ((m / a) ** 2 + (n / b) ** 2 + (l / d) ** 2)] * mult
    return np.sort(vals)[:count]


def test_box_cavity_eigenfrequencies(b2p_ctx):
    from palace_b200 import capi

    capi.set_stream(b2p_ctx)
    size = (1.0, 0.83, 0.61)
    p, nev = 3, 6
    prob = common.make_problem(n=(4, 3, 2), p=p, mesh_order=1, warp=0.0, scramble=7, n_attr=1, size=size)
    geom = common.gpu_geom(b2p_ctx, prob)
    orders = asm.p_sequence(p)
    nd = {q: hs.build_nd_space(prob.mesh, prob.topo, q) for q in orders}
    h1 = {q: hs.build_h1_space(prob.mesh, prob.topo, q) for q in orders}
    lam_exact = _analytic(size, nev)
    # target inside the physical spectrum: the nev eigenvalues nearest to sigma are all physical (the
    # curl-curl nullspace, lambda = 0, is further away than the sixth mode)
    sigma = 40.0
    assert np.abs(lam_exact - sigma).max() < 0.5 * sigma
    ident = cf.coeff_ctx()
    # system operator A = K - sigma M and the shifted positive preconditioner matrix K + sigma M
    # (GetSystemMatrix / GetPreconditionerMatrix with PCMatShifted, spaceoperator.cpp:945-1008)
    blob_A = cf.coeff_ctx_pair(cf.coeff_ctx(a=-sigma), cf.coeff_ctx(a=1.0))
    blob_P = cf.coeff_ctx_pair(cf.coeff_ctx(a=+sigma), cf.coeff_ctx(a=1.0))
    blob_G = cf.coeff_ctx(a=sigma)
    A = common.gpu_par_operator(b2p_ctx, geom, prob, O.CURLCURL_MASS, blob_A, space=nd[p])
    M = common.gpu_par_operator(b2p_ctx, geom, prob, O.ND_MASS, ident, space=nd[p])
    Pl, AG = {}, {}
    Pl[p] = common.gpu_par_operator(b2p_ctx, geom, prob, O.CURLCURL_MASS, blob_P, space=nd[p])
    AG[p] = common.gpu_par_operator(b2p_ctx, geom, prob, O.H1_DIFFUSION, blob_G, space=h1[p])
    for q in orders[:-1]:
        Pl[q] = common.gpu_par_operator(b2p_ctx, geom, prob, O.CURLCURL_MASS, blob_P, space=nd[q], fine_op=Pl[p].local_op)
        AG[q] = common.gpu_par_operator(b2p_ctx, geom, prob, O.H1_DIFFUSION, blob_G, space=h1[q], fine_op=AG[p].local_op)
    G = [common.gpu_interp(b2p_ctx, h1[q], nd[q], asm.gradient_comps(q)) for q in orders]
    P = [common.gpu_interp(b2p_ctx, nd[a], nd[b], asm.nd_prolongation_comps(a, b)) for a, b in zip(orders[:-1], orders[1:])]
    coarse = capi.Solver.krylov(b2p_ctx, capi.CG, rel_tol=1e-12, max_it=3000)
    cj = capi.Solver.jacobi(b2p_ctx)
    cj.set_operator(Pl[orders[0]])
    coarse.set_preconditioner(cj)
    coarse.set_operator(Pl[orders[0]])
    mg = capi.Solver.gmg(b2p_ctx, coarse, P, G, cycle_it=1, smooth_it=1, cheby_order=6)
    mg.gmg_set_operators([Pl[q] for q in orders], [AG[q] for q in orders])
    ksp = capi.Solver.krylov(b2p_ctx, capi.FGMRES, rel_tol=1e-12, max_it=200, max_dim=200)
    ksp.set_operator(A)
    ksp.set_preconditioner(mg)

    n = nd[p].ndofs
    free = np.setdiff1d(np.arange(n), nd[p].ess_dofs)
    xd = torch.zeros(n, dtype=torch.float64, device="cuda")
    yd = torch.zeros(n, dtype=torch.float64, device="cuda")
    its = []

    def to_full(v):
        f = np.zeros(n)
        f[free] = v
        return f

    def opinv(v):  # (K - sigma M)^-1 v on the free dofs, host vectors in and out like ARPACK's reverse communication
        xd.copy_(torch.from_numpy(to_full(v)))
        ksp.mult(xd, yd)
        its.append(ksp.stats()["its"])
        return yd.cpu().numpy()[free]

    def mmul(v):
        xd.copy_(torch.from_numpy(to_full(v)))
        M.mult(xd, yd)
        return yd.cpu().numpy()[free]

    nf = free.size
    Kop = spla.LinearOperator((nf, nf), matvec=lambda v: None, dtype=np.float64)  # unused in shift-invert mode
    lam_gpu = spla.eigsh(Kop, k=nev, M=spla.LinearOperator((nf, nf), matvec=mmul, dtype=np.float64), sigma=sigma, which="LM",
                         OPinv=spla.LinearOperator((nf, nf), matvec=opinv, dtype=np.float64), tol=1e-12,
                         v0=np.random.default_rng(0).standard_normal(nf), return_eigenvectors=False)
    lam_gpu = np.sort(lam_gpu)

    # the same eigenproblem on the oracle's assembled matrices
    Ko = common.oracle_matrix(prob, O.CURLCURL, ident, space=nd[p], eliminate=False)[free][:, free]
    Mo = common.oracle_matrix(prob, O.ND_MASS, ident, space=nd[p], eliminate=False)[free][:, free]
    lam_ref = np.sort(spla.eigsh(Ko.tocsc(), k=nev, M=Mo.tocsc(), sigma=sigma, which="LM", tol=1e-13, return_eigenvectors=False,
                                 v0=np.random.default_rng(1).standard_normal(nf)))

    rel_ref = np.abs(lam_gpu - lam_ref) / lam_ref
    rel_exact = np.abs(lam_gpu - lam_exact) / lam_exact
    print("eigenvalues k^2 (GPU)   :", lam_gpu)
    print("rel. err vs oracle      :", rel_ref, " FGMRES its/solve:", int(np.mean(its)))
    print("rel. err vs closed form :", rel_exact)
    # frequency ~ sqrt(lambda): relative frequency error is half the eigenvalue error
    assert rel_ref.max() < 1e-8
    assert rel_exact.max() < 2e-3
    assert np.mean(its) < 80  # indefinite shifted system (two modes below sigma), tol 1e-12
