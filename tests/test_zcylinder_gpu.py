"""GPU, end to end against the reference's own output: the cylinder-cavity example (curved HEX27 mesh, Nedelec order 4,
/root/reference/examples/cylinder/cavity_pec.json) solved through the whole device stack -- ARPACK shift-and-invert with
host vectors as Palace drives it (linalg/arpack.cpp:631-659) over the device FGMRES + p-multigrid (p = 1, 2, 4, Chebyshev +
Hiptmair smoothing) solve of (K - sigma M) x = M v -- must reproduce the eigenfrequencies stored by the reference's
regression suite (test/data/regression/ref/cylinder/cavity_pec/eig.csv; fixture tests/golden/cylinder_cavity_pec.npz) to the
north-star tolerance of 1e-8 relative (the reference's own regression tolerance is 1e-4)."""
import numpy as np
import pytest
import scipy.sparse.linalg as spla

from oracle import pyoracle as O
from palace_b200.host import assemble as asm
from palace_b200.host import coeff as cf
from palace_b200.host import hexspace as hs
from tests import common
from tests.test_cylinder_golden import FIX, _Mesh, frequencies_ghz, target_lambda

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def _problem(p):
    mesh = _Mesh(FIX)
    topo = hs.build_topology(mesh)
    q1d, mesh_order = p + 1, 2
    nodes = hs.gauss_lobatto(mesh_order + 1)
    qx, _ = hs.gauss_legendre(q1d)
    nB, nG = hs.lagrange_table(nodes, qx)
    xe = np.ascontiguousarray(FIX["xe2"])
    attr1 = np.ones(mesh.ne, dtype=np.int32)
    mesh.attr = attr1
    qd = O.geom_hex_qdata(xe, attr1, mesh_order, q1d)
    return common.Problem(mesh, topo, p, q1d, mesh_order, xe, hs.build_nd_space(mesh, topo, p), hs.build_h1_space(mesh, topo, p),
                          hs.tables_1d(p, q1d), (nB, nG), qd)


@pytest.mark.parametrize("nev,as_sum", [(6, False), (3, True)])
def test_cylinder_cavity_eigenfrequencies_match_the_reference(b2p_ctx, nev, as_sum):
    from palace_b200 import capi

    capi.set_stream(b2p_ctx)
    p = int(FIX["order"])
    prob = _problem(p)
    geom = common.gpu_geom(b2p_ctx, prob)
    orders = asm.p_sequence(p)                                      # LOGARITHMIC coarsening: 1, 2, 4 (multigrid.hpp:44-70)
    nd = {q: (prob.nd if q == p else hs.build_nd_space(prob.mesh, prob.topo, q)) for q in orders}
    h1 = {q: (prob.h1 if q == p else hs.build_h1_space(prob.mesh, prob.topo, q)) for q in orders}
    sigma = target_lambda()
    ident = cf.coeff_ctx()
    blob_A = cf.coeff_ctx_pair(cf.coeff_ctx(a=-sigma), cf.coeff_ctx(a=1.0))     # K - sigma M
    blob_P = cf.coeff_ctx_pair(cf.coeff_ctx(a=+sigma), cf.coeff_ctx(a=1.0))     # shifted positive preconditioner matrix
    blob_G = cf.coeff_ctx(a=sigma)
    M = common.gpu_par_operator(b2p_ctx, geom, prob, O.ND_MASS, ident, space=nd[p])
    Pl, AG = {}, {}
    if as_sum:
        # the way Palace builds them: K and M assembled once, system / preconditioner matrices as sums a0 K + a2 M
        # (BuildParSumOperator, rap.cpp:764-829) -- here each sum runs as one fused element operator
        Kop = common.gpu_op(b2p_ctx, geom, prob, O.CURLCURL, ident)
        Mop = common.gpu_op(b2p_ctx, geom, prob, O.ND_MASS, ident)
        A = capi.Operator.par(b2p_ctx, nd[p].ndofs, nd[p].ndofs, [Kop, Mop], [1.0, -sigma], nd[p].ess_dofs, diag_policy=1)
        assert A.is_fused()
        Psum = capi.Op.create_sum(b2p_ctx, [Kop, Mop], [1.0, sigma])
        Pl[p] = capi.Operator.par(b2p_ctx, nd[p].ndofs, nd[p].ndofs, [Psum], None, nd[p].ess_dofs, diag_policy=1)
        Pl[p].local_op = Psum
    else:
        A = common.gpu_par_operator(b2p_ctx, geom, prob, O.CURLCURL_MASS, blob_A, space=nd[p])
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
    ksp = capi.Solver.krylov(b2p_ctx, capi.FGMRES, rel_tol=1e-12, max_it=300, max_dim=300)
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

    def opinv(v):
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
    lam = spla.eigsh(Kop, k=nev, M=spla.LinearOperator((nf, nf), matvec=mmul, dtype=np.float64), sigma=sigma, which="LA",
                     OPinv=spla.LinearOperator((nf, nf), matvec=opinv, dtype=np.float64), tol=1e-11,
                     v0=np.random.default_rng(0).standard_normal(nf), return_eigenvectors=False)
    f = frequencies_ghz(np.sort(lam))
    ref = FIX["ref_f_re_ghz"][:nev]
    rel = np.abs(f.real - ref) / ref
    print("Re f (GHz), device stack:", f.real)
    print("rel. error vs the reference's eig.csv:", rel, " FGMRES its/solve:", int(np.mean(its)))
    assert rel.max() < 1e-8
    assert (np.abs(f.imag - FIX["ref_f_im_ghz"][:nev]) / FIX["ref_f_im_ghz"][:nev]).max() < 1e-6
