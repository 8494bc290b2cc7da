"""CPU: the host layer's non-conforming hexahedral mesh (one level of hanging faces) and the conforming prolongation P of its ND
space (palace_b200/host/nonconforming.py) -- the matrix the reference gets from mfem::ParFiniteElementSpace on AMR meshes and
multiplies by around the local operator (/root/reference/palace/linalg/rap.cpp:195-234). MFEM is not in /root/reference, so P is
checked by what it must achieve: tangential continuity across the hanging faces for every true-dof vector, and cavity eigenvalues
of the constrained curl-curl problem that converge to the closed form without spurious modes."""
import numpy as np
import pytest
import scipy.linalg as sla
import scipy.sparse as sparse

from oracle import pyoracle as O
from palace_b200.host import hexspace as hs
from palace_b200.host import nonconforming as nc
from tests import common


@pytest.mark.parametrize("p,scramble", [(1, None), (2, 5), (3, 11)])
def test_prolongation_gives_tangentially_continuous_fields(p, scramble):
    hb = nc.hanging_box_mesh(nc=(1, 2, 1), nfx=1, h=1.0, scramble_seed=scramble)
    cs = nc.build_constrained_nd_space(hb, p)
    sp_, mesh = cs.space, hb.mesh
    n_true = cs.P.shape[1]
    assert cs.slaves.size > 0 and n_true + cs.slaves.size == sp_.ndofs
    # true-dof rows are unit rows
    Pt = cs.P[np.nonzero(cs.true_of >= 0)[0]]
    assert abs(Pt - sparse.identity(n_true)).max() == 0.0
    # slaves per coarse interface face: 12 fine edges in the plane (p dofs each) + 4 fine faces (2 p (p - 1) each), shared
    # edges between neighbouring coarse faces counted once: 2 coarse faces side by side share 2 fine edges
    assert cs.slaves.size == 2 * (12 * p + 4 * 2 * p * (p - 1)) - 2 * p
    rng = np.random.default_rng(0)
    xL = cs.P @ (rng.random(n_true) - 0.5)

    def field(e, xp):
        x0, J = nc._affine(mesh, e)
        xi = np.linalg.solve(J, xp - x0)
        uhat = nc._nd_basis_at(p, np.clip(xi, 0.0, 1.0)).T @ (sp_.lex_sign[e] * xL[sp_.lex_gid[e]])
        return np.linalg.solve(J.T, uhat)  # u = J^-T u^

    worst = 0.0
    for ef in np.nonzero(~hb.coarse)[0]:
        v = mesh.verts[mesh.elems[ef]]
        if abs(v[:, 0].min() - hb.x_interface) > 1e-12:
            continue
        lo, hi = v.min(axis=0), v.max(axis=0)
        for _ in range(5):
            xp = np.array([hb.x_interface, *(lo[1:] + rng.random(2) * (hi[1:] - lo[1:]))])
            ec = next(e for e in np.nonzero(hb.coarse)[0]
                      if np.all(xp >= mesh.verts[mesh.elems[e]].min(axis=0) - 1e-12) and np.all(xp <= mesh.verts[mesh.elems[e]].max(axis=0) + 1e-12))
            uf, uc = field(ef, xp), field(ec, xp)
            worst = max(worst, abs(uf[1] - uc[1]), abs(uf[2] - uc[2]))
    assert worst < 1e-12


def test_constrained_cavity_eigenvalues_match_the_closed_form():
    """PEC box 2 x 1 x 1, coarse cell + 8 half-size cells, order 4: k^2 = pi^2 (l^2 / 4 + m^2 + n^2), two indices non-zero
    (orders 2 / 3 / 4 give 9.5e-3 / 6.4e-3 / 7.3e-6 on this 9-element mesh)."""
    p = 4
    hb = nc.hanging_box_mesh(nc=(1, 1, 1), nfx=2, h=1.0, scramble_seed=3)
    cs = nc.build_constrained_nd_space(hb, p)
    prob = common.problem_on_mesh(hb.mesh, p)
    assert prob.nd.ndofs == cs.space.ndofs and np.array_equal(prob.nd.lex_gid, cs.space.lex_gid)
    K = common.oracle_matrix(prob, O.CURLCURL, common.coefficient(O.CURLCURL, 1, "const"), eliminate=False)
    M = common.oracle_matrix(prob, O.ND_MASS, common.coefficient(O.ND_MASS, 1, "const"), eliminate=False)
    Kt, Mt = (cs.P.T @ K @ cs.P).toarray(), (cs.P.T @ M @ cs.P).toarray()
    free = np.setdiff1d(np.arange(cs.P.shape[1]), cs.ess_tdofs)
    w = sla.eigh(Kt[np.ix_(free, free)], Mt[np.ix_(free, free)], eigvals_only=True)
    w = w[w > 1.0]  # the gradient null space sits at ~1e-12
    exact = sorted(np.pi ** 2 * (l * l / 4.0 + m * m + n * n) for l in range(4) for m in range(3) for n in range(3)
                   if (l > 0) + (m > 0) + (n > 0) >= 2 for _ in range(2 if l * m * n > 0 else 1))[:6]
    assert np.allclose(w[:6], exact, rtol=5e-5), (w[:6], exact)
    # no spurious modes between the null space and the first cavity mode
    assert w[0] > 0.9999 * exact[0]


@pytest.mark.parametrize("p", [1, 2, 3])
def test_constrained_gradients_are_constrained_fields(p):
    """The discrete de Rham property the Hiptmair smoother relies on survives the constraints: G (P_h1 phi) = P_nd (R_nd G P_h1 phi)
    for every true-dof potential, i.e. gradients of continuous potentials are tangentially continuous ND fields, and the
    constrained curl-curl matrix annihilates them."""
    hb = nc.hanging_box_mesh(nc=(1, 2, 1), nfx=2, h=1.0, scramble_seed=p)
    cn, ch = nc.build_constrained_nd_space(hb, p), nc.build_constrained_h1_space(hb, p)
    assert np.array_equal(ch.P.sum(axis=1).A1.round(12), np.ones(ch.space.ndofs))  # partition of unity: constants are reproduced
    G = common.oracle_interp(ch.space, cn.space, hs.discrete_gradient_matrix(p))
    R = nc.restriction_matrix(cn)
    phi = np.random.default_rng(0).random(ch.P.shape[1])
    gL = G @ (ch.P @ phi)
    assert np.abs(cn.P @ (R @ gL) - gL).max() < 1e-12 * np.abs(gL).max()
    prob = common.problem_on_mesh(hb.mesh, p)
    K = common.oracle_matrix(prob, O.CURLCURL, common.coefficient(O.CURLCURL, 1, "const"), eliminate=False)
    Kt = cn.P.T @ K @ cn.P
    assert np.abs(Kt @ (R @ gL)).max() < 1e-11 * abs(Kt).max() * np.abs(gL).max()
