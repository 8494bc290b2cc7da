"""GPU parity of the two mixed curl integrators on one ND space -- the Floquet-periodic terms of the reference
(/root/reference/palace/models/spaceoperator.cpp:305-309):

  B2P_ND_WEAKCURL   MixedVectorWeakCurlIntegrator  -(F u, curl v)    f_apply_hcurlhdiv_33   integ/mixedveccurl.cpp:68-117
  B2P_ND_MIXEDCURL  MixedVectorCurlIntegrator       (F^T curl u, v)  f_apply_hdivhcurl_33   integ/mixedveccurl.cpp:23-66

through the C ABI (dense-basis operators, b2p_op_create_dense) against the oracle, whose pointwise arithmetic for these kinds is
pinned to the reference's own hcurlhdiv_33_qf.h compiled in place (tests/test_oracle_golden.py). Coefficient contexts are built
as the two integrators build theirs: a = -1 for the weak curl, transposed matrices for the curl."""
import numpy as np
import pytest

from oracle import pyoracle as O
from oracle import solvers as S
from palace_b200.host import coeff as cf
from tests import common

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu
RTOL = 1e-12


def _dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def _rel(a, b):
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


def periodic_ctx(kind, n_attr, seed=5):
    """Non-symmetric per-material matrices (the Floquet coefficient is a cross-product matrix times mu^-1, materialoperator.cpp)."""
    am, mc = cf.test_suite_coefficient(n_attr, "matrix")
    mc = mc + 0.3 * (np.random.default_rng(seed).random(mc.shape) - 0.5)
    if kind == O.ND_WEAKCURL:
        return cf.coeff_ctx(am, mc, a=-1.0)                  # PopulateCoefficientContext(space_dim, Q, transpose, -1.0)
    return cf.coeff_ctx(am, mc, a=1.0, transpose=True)       # AddDomainIntegrator<MixedVectorCurlIntegrator>(*fp, true)


def dense_op(ctx, prob, kind, blob, curl_orient=None):
    from palace_b200 import capi

    geom = capi.Geom.general(ctx, prob.qdata_ref)
    sp = prob.nd
    interp, curl, _ = O.nd_hex_tables(sp.p, prob.q1d)
    idx, ori = sp.native_restriction()
    return capi.Op.create_dense(ctx, geom, kind, sp.ndofs, idx, ori, interp, curl, blob, curl_orient=curl_orient)


@pytest.mark.parametrize("kind", [O.ND_WEAKCURL, O.ND_MIXEDCURL])
@pytest.mark.parametrize("p", [1, 2, 3])
def test_mixed_curl_apply_and_diagonal_match_oracle(b2p_ctx, p, kind):
    prob = common.make_problem(n=(3, 3, 2), p=p, n_attr=3)  # 18 warped elements: two full batches of 8 and a ragged one
    blob = periodic_ctx(kind, 3)
    op = dense_op(b2p_ctx, prob, kind, blob)
    n = prob.nd.ndofs
    x = np.random.default_rng(1).random(n) - 0.5
    y_ref = common.oracle_apply(prob, kind, blob, x)
    yd = torch.full((n,), 3.0, dtype=torch.float64, device="cuda")
    op.apply(_dev(x), yd)
    assert _rel(yd.cpu().numpy(), y_ref) < RTOL
    op.apply_add_ex(-0.5, _dev(x), yd)
    assert _rel(yd.cpu().numpy(), 0.5 * y_ref) < RTOL
    dd = torch.zeros(n, dtype=torch.float64, device="cuda")
    op.diag_add(dd)
    d_ref = common.oracle_diag(prob, kind, blob)
    assert np.abs(dd.cpu().numpy() - d_ref).max() < RTOL * np.abs(d_ref).max()


@pytest.mark.parametrize("kind", [O.ND_WEAKCURL, O.ND_MIXEDCURL])
def test_transposed_apply_is_the_other_kind_with_the_transposed_coefficient(b2p_ctx, kind):
    prob = common.make_problem(n=(2, 3, 2), p=2, n_attr=3)
    blob = periodic_ctx(kind, 3)
    op = dense_op(b2p_ctx, prob, kind, blob)
    n = prob.nd.ndofs
    rng = np.random.default_rng(2)
    x, z = rng.random(n) - 0.5, rng.random(n) - 0.5
    A = common.oracle_matrix(prob, kind, blob, eliminate=False)
    yt = torch.zeros(n, dtype=torch.float64, device="cuda")
    op.apply_add_ex(1.0, _dev(z), yt, transpose=True)
    assert _rel(yt.cpu().numpy(), A.T @ z) < RTOL
    y = torch.zeros(n, dtype=torch.float64, device="cuda")
    op.apply_add_ex(1.0, _dev(x), y)
    assert abs(z @ y.cpu().numpy() - x @ yt.cpu().numpy()) < 1e-12 * np.linalg.norm(A @ x) * np.linalg.norm(z)


def test_floquet_pair_is_skew_symmetric_and_eliminates_essential_dofs(b2p_ctx):
    """-(F u, curl v) + (F^T curl u, v) as one ParOperator (two terms, applied one after the other: dense operators do not fuse):
    Mult against the oracle's assembled, eliminated sum; the pair is skew-symmetric, so MultTranspose = -Mult off the essential rows."""
    from palace_b200 import capi

    prob = common.make_problem(n=(3, 2, 2), p=2, n_attr=3)
    sp = prob.nd
    n = sp.ndofs
    bw, bc = periodic_ctx(O.ND_WEAKCURL, 3), periodic_ctx(O.ND_MIXEDCURL, 3)
    ow, oc = dense_op(b2p_ctx, prob, O.ND_WEAKCURL, bw), dense_op(b2p_ctx, prob, O.ND_MIXEDCURL, bc)
    Aw = common.oracle_matrix(prob, O.ND_WEAKCURL, bw, eliminate=False)
    Ac = common.oracle_matrix(prob, O.ND_MIXEDCURL, bc, eliminate=False)
    Asum = (0.7 * Aw + 0.7 * Ac).tocsr()
    assert abs(Asum + Asum.T).max() < 1e-12 * abs(Asum).max()
    A = capi.Operator.par(b2p_ctx, n, n, [ow, oc], [0.7, 0.7], sp.ess_dofs, diag_policy=1)
    assert not A.is_fused()
    Ael = S.eliminate(Asum, sp.ess_dofs)
    x = np.random.default_rng(3).random(n) - 0.5
    y = torch.empty(n, dtype=torch.float64, device="cuda")
    A.mult(_dev(x), y)
    assert _rel(y.cpu().numpy(), Ael @ x) < RTOL
    yt = torch.empty(n, dtype=torch.float64, device="cuda")
    A.mult_transpose(_dev(x), yt)
    assert _rel(yt.cpu().numpy(), Ael.T @ x) < RTOL
    y2 = _dev(np.ones(n))
    A.add_mult(_dev(x), y2, -2.0)
    assert _rel(y2.cpu().numpy(), 1.0 - 2.0 * (Ael @ x)) < RTOL


def test_mixed_curl_with_curl_oriented_restriction(b2p_ctx):
    """The tridiagonal restriction of ND tets / prisms (restriction.cpp:301-368) around the mixed kinds, synthetic transformations."""
    prob = common.make_problem(n=(3, 2, 2), p=2, n_attr=2)
    sp = prob.nd
    P, ne = sp.P, prob.mesh.ne
    rng = np.random.default_rng(4)
    co = np.zeros((ne, P, 3), dtype=np.int8)
    co[:, :, 1] = rng.choice([-1, 1], size=(ne, P))
    for e in range(ne):
        for j in rng.choice(np.arange(0, P - 1, 2), size=P // 6, replace=False):
            blk = rng.integers(-1, 2, size=(2, 2))
            co[e, j, 1], co[e, j, 2] = blk[0, 0], blk[0, 1]
            co[e, j + 1, 0], co[e, j + 1, 1] = blk[1, 0], blk[1, 1]
    idx, _ = sp.native_restriction()
    interp, curl, _ = O.nd_hex_tables(sp.p, prob.q1d)
    x = rng.random(sp.ndofs) - 0.5
    for kind in (O.ND_WEAKCURL, O.ND_MIXEDCURL):
        blob = periodic_ctx(kind, 2)
        op = dense_op(b2p_ctx, prob, kind, blob, curl_orient=co)
        y_ref = O.apply_add_co(kind, interp, curl, idx, co, prob.qdata_ref, blob, x, np.zeros(sp.ndofs))
        yd = torch.empty(sp.ndofs, dtype=torch.float64, device="cuda")
        op.apply(_dev(x), yd)
        assert _rel(yd.cpu().numpy(), y_ref) < RTOL


def test_floquet_terms_in_a_complex_operator(b2p_ctx):
    """K + i (weak curl + curl) - omega^2 M on split complex vectors (ComplexParOperator over the term list; the periodic pair
    carries the imaginary unit, so the matrix is Hermitian when F is real and the pair skew-symmetric)."""
    from palace_b200 import capi

    prob = common.make_problem(n=(2, 2, 2), p=2, n_attr=2)
    sp = prob.nd
    n = sp.ndofs
    bk = common.coefficient(O.CURLCURL, 2, "matrix")
    bm = common.coefficient(O.ND_MASS, 2, "matrix")
    bw, bc = periodic_ctx(O.ND_WEAKCURL, 2), periodic_ctx(O.ND_MIXEDCURL, 2)
    kinds = [O.CURLCURL, O.ND_MASS, O.ND_WEAKCURL, O.ND_MIXEDCURL]
    blobs = [bk, bm, bw, bc]
    ops = [dense_op(b2p_ctx, prob, k, b) for k, b in zip(kinds, blobs)]
    cr, ci = [1.0, -2.3, 0.0, 0.0], [0.0, 0.1, 0.9, 0.9]
    mats = [common.oracle_matrix(prob, k, b, eliminate=False) for k, b in zip(kinds, blobs)]
    Z = sum((a + 1j * b) * M for a, b, M in zip(cr, ci, mats)).tocsr()
    A = capi.ComplexOperator.par(b2p_ctx, n, n, ops, [a + 1j * b for a, b in zip(cr, ci)], None, diag_policy=1)
    rng = np.random.default_rng(6)
    x = rng.random(n) - 0.5 + 1j * (rng.random(n) - 0.5)
    yr, yi = torch.empty(n, dtype=torch.float64, device="cuda"), torch.empty(n, dtype=torch.float64, device="cuda")
    A.mult(_dev(x.real), _dev(x.imag), yr, yi)
    y = yr.cpu().numpy() + 1j * yi.cpu().numpy()
    assert _rel(y, Z @ x) < RTOL


@pytest.mark.parametrize("p", [2, 3])
def test_mixed_curl_kinds_on_scrambled_tetrahedra(b2p_ctx, p):
    """Real simplex tables, the tridiagonal curl-oriented restriction of scrambled ND tets (restriction.cpp:301-368), curved
    (quadratic) geometry: both kinds against the oracle, and the integration-by-parts identity the Floquet terms rest on --
    with F constant and u, v vanishing tangentially on the boundary, (F^T curl u, v) = (curl u, F v) and -(F u, curl v) are
    transposes of each other up to sign: x^T (A_weak(F) + A_curl(F^T-context)) x = 0 for every x."""
    from palace_b200 import capi
    from palace_b200.host import tetspace as ts

    mesh = ts.box_tet_mesh((2, 2, 1), (1.0, 0.8, 0.9), jitter=0.25, scramble_seed=5, n_attr=2, warp_amp=0.03)
    sp = ts.build_nd_tet_space(mesh, p)
    interp, curl, qpts, qw = ts.nd_tet_tables(p)
    qd = ts.geom_qdata(mesh.node_coords(2), mesh.attr, 2, qpts, qw)
    geom = capi.Geom.general(b2p_ctx, qd)
    rng = np.random.default_rng(8)
    x = rng.random(sp.ndofs) - 0.5
    ys = {}
    for kind in (O.ND_WEAKCURL, O.ND_MIXEDCURL):
        blob = periodic_ctx(kind, 2)
        op = capi.Op.create_dense(b2p_ctx, geom, kind, sp.ndofs, sp.idx, None, interp, curl, blob, curl_orient=sp.curl_orient)
        y_ref = O.apply_add_co(kind, interp, curl, sp.idx, sp.curl_orient, qd, blob, x, np.zeros(sp.ndofs))
        yd = torch.empty(sp.ndofs, dtype=torch.float64, device="cuda")
        op.apply(_dev(x), yd)
        assert _rel(yd.cpu().numpy(), y_ref) < RTOL
        ys[kind] = yd.cpu().numpy()
    # the two contexts of periodic_ctx hold -F and F^T of the SAME F: the pair is skew-symmetric
    assert abs(x @ (ys[O.ND_WEAKCURL] + ys[O.ND_MIXEDCURL])) < 1e-12 * np.linalg.norm(x) * np.linalg.norm(ys[O.ND_WEAKCURL])
