"""DivFreeSolver (/root/reference/palace/linalg/divfree.cpp:42-186) through the C ABI: y <- y + G psi with
(G^T M_eps G) psi = WeakDiv y on the H1 p-hierarchy.

The reference has no unit test for it, so the oracle is a restatement on explicitly assembled matrices:
  * the weak divergence the way the reference assembles it -- MixedVectorWeakDivergenceIntegrator
    (fem/integ/mixedvecgrad.cpp:148-208): ND interpolation -> the H(curl) mass quadrature function with the coefficient negated
    -> H1 gradient -- element by element in NumPy from the oracle's basis tables and pointwise D;
  * the H1 diffusion matrix with the same coefficient, essential rows eliminated the ParOperator way (DIAG_ONE);
  * a sparse direct solve instead of the PCG.
The library computes the weak divergence as -G^T (M_eps y) (one ND mass apply + the transposed discrete gradient); the first test
pins that identity on the oracle side."""
import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from oracle import pyoracle as O
from oracle import solvers as S
from palace_b200.host import assemble as asm
from palace_b200.host import hexspace as hs
from tests import common

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def _dev(a):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).cuda()


def _rel(a, b):
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


def oracle_weak_divergence(prob, nd, h1, blob):
    """Sparse [n_h1 x n_nd] matrix of -(eps u, grad v), assembled the reference's way (trial: ND Interp, test: H1 Grad, the
    H(curl) mass quadrature function scaled by -1: mixedvecgrad.cpp:199-207)."""
    q1d = prob.q1d
    interp, _, _ = O.nd_hex_tables(nd.p, q1d)       # [3][Q][P_nd], native dof order
    _, grad, _ = O.h1_hex_tables(h1.p, q1d)         # [3][Q][P_h1], lexicographic
    idx, ori = nd.native_restriction()
    Pn, Ph = interp.shape[2], grad.shape[2]
    Q = interp.shape[1]
    rows, cols, vals = [], [], []
    zero = np.zeros((3, Q))
    for e in range(idx.shape[0]):
        We = np.empty((Ph, Pn))
        for b in range(Pn):
            u = np.ascontiguousarray(interp[:, :, b])
            v, _ = O.apply_D(O.ND_MASS, blob, prob.qdata_ref[e], u, zero)   # D u at the quadrature points
            We[:, b] = -np.einsum("cqa,cq->a", grad, v)
        sgn = ori[e].astype(np.float64)  # u_nat = orient * x[idx] (restriction.cpp:281-297)
        rows.append(np.repeat(h1.lex_gid[e], Pn))
        cols.append(np.tile(idx[e], Ph))
        vals.append((We * sgn[None, :]).ravel())
    return sp.coo_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(h1.ndofs, nd.ndofs)).tocsr()


@pytest.fixture(scope="module", params=[2, 3])
def setup(request, b2p_ctx):
    from palace_b200 import capi

    p = request.param
    prob = common.make_problem(n=(3, 2, 2), p=p)
    geom = common.gpu_geom(b2p_ctx, prob)
    orders = asm.p_sequence(p)
    blob = common.coefficient(O.ND_MASS, 3, "matrix", a_mass=1.3)       # the permittivity: same blob for both operators
    nd = prob.nd
    h1 = {q: (prob.h1 if q == p else hs.build_h1_space(prob.mesh, prob.topo, q)) for q in orders}
    # ND mass WITHOUT essential dofs, discrete gradient, H1 hierarchy (diffusion with eps, essential dofs DIAG_ONE)
    mop = common.gpu_op(b2p_ctx, geom, prob, O.ND_MASS, blob)
    Mnd = capi.Operator.par(b2p_ctx, nd.ndofs, nd.ndofs, [mop], None, None, diag_policy=1)
    G = common.gpu_interp(b2p_ctx, h1[p], nd, asm.gradient_comps(p))
    fineK = common.gpu_par_operator(b2p_ctx, geom, prob, O.H1_DIFFUSION, blob, space=h1[p])
    K = [fineK if q == p else common.gpu_par_operator(b2p_ctx, geom, prob, O.H1_DIFFUSION, blob, space=h1[q], fine_op=fineK.local_op)
         for q in orders]
    P = [common.gpu_interp(b2p_ctx, h1[a], h1[b], asm.h1_prolongation_comps(a, b)) for a, b in zip(orders[:-1], orders[1:])]
    # oracle matrices
    Mo = common.oracle_matrix(prob, O.ND_MASS, blob, space=nd, eliminate=False)
    Go = common.oracle_interp(h1[p], nd, hs.discrete_gradient_matrix(p))
    Ko_full = common.oracle_matrix(prob, O.H1_DIFFUSION, blob, space=h1[p], eliminate=False)
    Ko = S.eliminate(Ko_full, h1[p].ess_dofs)
    return dict(p=p, prob=prob, nd=nd, h1=h1, orders=orders, blob=blob, Mnd=Mnd, G=G, K=K, P=P, Mo=Mo, Go=Go, Ko=Ko, Ko_full=Ko_full,
                keep=[mop, geom])


def oracle_project(s, y):
    ess = s["h1"][s["p"]].ess_dofs
    rhs = -(s["Go"].T @ (s["Mo"] @ y))
    rhs[ess] = 0.0
    psi = spla.spsolve(s["Ko"].tocsc(), rhs)
    return y + s["Go"] @ psi, psi


def test_weak_divergence_is_minus_gradient_transpose_times_mass(setup):
    """The operator the reference partially assembles for WeakDiv equals -G^T M_eps (and G^T M_eps G the H1 diffusion matrix)."""
    s = setup
    W = oracle_weak_divergence(s["prob"], s["nd"], s["h1"][s["p"]], s["blob"])
    W2 = -(s["Go"].T @ s["Mo"])
    assert abs(W - W2).max() < 1e-12 * abs(W2).max()
    K2 = s["Go"].T @ s["Mo"] @ s["Go"]
    assert abs(K2 - s["Ko_full"]).max() < 1e-12 * abs(s["Ko_full"]).max()


@pytest.mark.parametrize("levels", ["multigrid", "one_level"])
def test_projection_matches_the_oracle(b2p_ctx, setup, levels):
    from palace_b200 import capi

    s = setup
    p, h1p = s["p"], s["h1"][s["p"]]
    if levels == "multigrid":
        dv = capi.DivFree(b2p_ctx, s["Mnd"], s["G"], s["K"], s["P"], h1p.ess_dofs, p, tol=1e-11, max_it=200, coarse_type=1, coarse_tol=1e-3)
    else:
        dv = capi.DivFree(b2p_ctx, s["Mnd"], s["G"], s["K"][-1:], [], h1p.ess_dofs, p, tol=1e-11, max_it=2000, coarse_type=0)
    rng = np.random.default_rng(41)
    y = rng.standard_normal(s["nd"].ndofs)
    y_ref, _ = oracle_project(s, y)
    yd = _dev(y)
    dv.mult(yd)
    st = dv.stats()
    assert st["converged"] and st["num_mult"] == 1 and st["its"] == st["num_mult_its"]
    if levels == "multigrid":
        assert st["its"] <= 25, st  # p-multigrid + Chebyshev on an SPD system: a few iterations per digit
    out = yd.cpu().numpy()
    assert _rel(out, y_ref) < 1e-9
    # discretely divergence-free on the free H1 dofs, the irrotational part went away, projecting again is the identity
    free = np.setdiff1d(np.arange(h1p.ndofs), h1p.ess_dofs)
    div0 = (s["Go"].T @ (s["Mo"] @ y))[free]
    div1 = (s["Go"].T @ (s["Mo"] @ out))[free]
    assert np.linalg.norm(div1) < 1e-9 * np.linalg.norm(div0)
    yd2 = _dev(out)
    dv.mult(yd2)
    assert _rel(yd2.cpu().numpy(), out) < 1e-9
    # a pure gradient of a potential that vanishes on the essential dofs is removed completely
    phi = rng.standard_normal(h1p.ndofs)
    phi[h1p.ess_dofs] = 0.0
    yg = _dev(s["Go"] @ phi)
    dv.mult(yg)
    assert np.linalg.norm(yg.cpu().numpy()) < 1e-8 * np.linalg.norm(s["Go"] @ phi)


def test_complex_projection_is_the_joint_solve_of_both_parts(b2p_ctx, setup):
    """DivFreeSolver<ComplexVector>: one CG iteration on the complex vector (ComplexParOperator(M, nullptr)) = CG on the stacked
    parts; the result is the projection of each part."""
    from palace_b200 import capi

    s = setup
    p, h1p = s["p"], s["h1"][s["p"]]
    dv = capi.DivFree(b2p_ctx, s["Mnd"], s["G"], s["K"], s["P"], h1p.ess_dofs, p, tol=1e-11, max_it=200, coarse_type=1, coarse_tol=1e-3)
    rng = np.random.default_rng(43)
    yr, yi = rng.standard_normal(s["nd"].ndofs), 1e-3 * rng.standard_normal(s["nd"].ndofs)  # parts of very different size
    ref_r, _ = oracle_project(s, yr)
    ref_i, _ = oracle_project(s, yi)
    dr, di = _dev(yr), _dev(yi)
    dv.mult_complex(dr, di)
    st = dv.stats()
    assert st["converged"] and st["its"] <= 25
    assert _rel(dr.cpu().numpy(), ref_r) < 1e-9
    # the joint residual criterion is relative to the norm of the WHOLE complex right-hand side: the small part is resolved to
    # the same absolute level as the large one
    assert np.linalg.norm(di.cpu().numpy() - ref_i) < 1e-9 * np.linalg.norm(ref_r)


def test_size_mismatch_is_reported(b2p_ctx, setup):
    from palace_b200 import capi

    s = setup
    with pytest.raises(capi.B2PError, match="do not match"):
        capi.DivFree(b2p_ctx, s["K"][-1], s["G"], s["K"], s["P"], s["h1"][s["p"]].ess_dofs, s["p"])
