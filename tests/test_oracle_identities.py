"""CPU: structural identities that pin the restated element conventions (MFEM's ND/H1 hex
elements are un-vendored, so no stored vectors exist at that boundary -- SURVEY §8c):
two independent transcriptions of GetDofMap() agree; curl(grad) = 0 discretely on meshes with
rotated element frames and curved geometry; G^T M G equals the H1 diffusion operator; symmetry;
mass of a constant field = volume; the box-cavity Rayleigh quotient."""
import numpy as np
import pytest

from oracle import pyoracle as O
from palace_b200.host import coeff as cf
from palace_b200.host import hexspace as hs
from tests import common


@pytest.mark.parametrize("p", [1, 2, 3, 4, 5, 6])
def test_dofmap_transcriptions_agree(p):
    a, b = O.nd_hex_dofmap(p), hs.nd_hex_dofmap(p)
    assert np.array_equal(a, b)
    nat = np.where(b < 0, -1 - b, b)
    assert sorted(nat) == list(range(3 * p * (p + 1) ** 2))


@pytest.mark.parametrize("p", [1, 2, 3, 4, 6])
def test_1d_tables_agree_with_long_double_oracle(p):
    t = hs.tables_1d(p)
    Bo, Bc, Gc, qw = O.nd_hex_1d(p, p + 1)
    assert np.abs(t.Bo - Bo).max() < 5e-15 and np.abs(t.Bc - Bc).max() < 5e-15
    assert np.abs(t.Gc - Gc).max() < 5e-13 and np.abs(t.qw - qw).max() < 1e-15
    assert np.allclose(Bo.sum(1), 1) and np.allclose(Bc.sum(1), 1) and np.abs(Gc.sum(1)).max() < 1e-12


def _grad(prob, phi):
    G = hs.discrete_gradient_matrix(prob.p)
    nd, h1 = prob.nd, prob.h1
    ue = np.einsum("lm,em->el", G, phi[h1.lex_gid]) * nd.lex_sign
    u = np.zeros(nd.ndofs)
    u[nd.lex_gid.ravel()] = ue.ravel()
    acc = np.zeros(nd.ndofs)
    np.add.at(acc, nd.lex_gid.ravel(), ue.ravel())
    assert np.abs(acc / nd.mult - u).max() < 1e-12  # every element computes the same value for shared dofs
    return u


@pytest.mark.parametrize("p", [1, 2, 3])
def test_curl_of_gradient_vanishes_and_GtMG_is_diffusion(p):
    prob = common.make_problem(p=p)
    ident = cf.coeff_ctx()
    phi = np.random.default_rng(0).random(prob.h1.ndofs)
    u = _grad(prob, phi)
    Ku = common.oracle_apply(prob, O.CURLCURL, ident, u)
    assert np.abs(Ku).max() < 1e-11 * np.abs(u).max()
    blob = common.coefficient(O.ND_MASS, 3, "matrix")
    Mu = common.oracle_apply(prob, O.ND_MASS, blob, u)
    Aphi = common.oracle_apply(prob, O.H1_DIFFUSION, blob, phi)
    assert abs(u @ Mu - phi @ Aphi) < 1e-12 * abs(phi @ Aphi)


def test_operator_is_symmetric_and_mass_of_constant_field_is_volume():
    prob = common.make_problem(p=2)
    ident = cf.coeff_ctx_pair(cf.coeff_ctx(), cf.coeff_ctx())
    rng = np.random.default_rng(1)
    x, z = rng.random(prob.nd.ndofs), rng.random(prob.nd.ndofs)
    Ax = common.oracle_apply(prob, O.CURLCURL_MASS, ident, x)
    Az = common.oracle_apply(prob, O.CURLCURL_MASS, ident, z)
    assert abs(z @ Ax - x @ Az) < 1e-12 * abs(z @ Ax)
    # phi = x-coordinate is in the (isoparametric, order >= mesh order) H1 space: grad phi = e_x
    h1 = prob.h1
    nodes = hs.gauss_lobatto(prob.p + 1)
    xe = prob.mesh.node_coords(prob.p, nodes)  # H1 nodes of order p
    phi = np.zeros(h1.ndofs)
    phi[h1.lex_gid.ravel()] = xe[:, 0, :].ravel()
    u = _grad(prob, phi)
    Mu = common.oracle_apply(prob, O.ND_MASS, cf.coeff_ctx(), u)
    vol = prob.qdata_ref[:, 1, :].sum()
    assert abs(vol - 1.0 * 0.7 * 0.9) < 1e-12
    assert abs(u @ Mu - vol) < 1e-11


def test_scrambled_element_frames_give_the_same_operator():
    """Rotating element-local frames changes idx/orient/qdata but not the global operator spectrum."""
    import scipy.sparse as sp
    import scipy.sparse.linalg  # noqa: F401

    def spectrum(scramble):
        prob = common.make_problem(n=(2, 2, 1), p=2, scramble=scramble, warp=0.03, n_attr=1)
        ident = cf.coeff_ctx_pair(cf.coeff_ctx(), cf.coeff_ctx())
        interp, curl, _ = O.nd_hex_tables(2, 3)
        idx, ori = prob.nd.native_restriction()
        Ae = O.element_matrices(O.CURLCURL_MASS, interp, curl, ori, prob.qdata_ref, ident, prob.nd.P)
        rows = np.repeat(idx, prob.nd.P, axis=1).ravel()
        cols = np.tile(idx, (1, prob.nd.P)).ravel()
        A = sp.coo_matrix((Ae.ravel(), (rows, cols)), shape=(prob.nd.ndofs,) * 2).toarray()
        return np.sort(np.linalg.eigvalsh(A))

    a, b = spectrum(None), spectrum(11)
    assert np.abs(a - b).max() < 1e-10 * np.abs(a).max()


@pytest.mark.parametrize("kind", [O.CURLCURL, O.ND_MASS, O.CURLCURL_MASS])
def test_blocked_reference_arm_equals_the_plain_oracle_apply(kind):
    """bench.py's reference arm (blocks of 8 elements, pinned thread pool, transposed restriction) performs the same
    arithmetic per element as the plain oracle apply -- including a ragged last block and signs."""
    from tests import common

    prob = common.make_problem(n=(3, 2, 3), p=2, scramble=3, warp=0.04, n_attr=2)  # 18 elements: blocks of 8, 8, 2
    blob = common.coefficient(kind, 2, "matrix")
    interp, curl, _ = O.nd_hex_tables(prob.nd.p, prob.q1d)
    idx, ori = prob.nd.native_restriction()
    x = np.random.default_rng(5).standard_normal(prob.nd.ndofs)
    y_ref = common.oracle_apply(prob, kind, blob, x)
    arm = O.BlockedApply(kind, interp, curl, np.ascontiguousarray(idx), np.ascontiguousarray(ori), prob.qdata_ref, blob, prob.nd.ndofs,
                         nthreads=3)
    y = np.zeros(prob.nd.ndofs)
    arm.apply_add(x, y)
    arm.apply_add(x, y)  # the pool is persistent: a second call accumulates again
    arm.close()
    assert np.linalg.norm(y - 2 * y_ref) <= 1e-13 * np.linalg.norm(y_ref)
    assert O.physical_cores() >= 1
