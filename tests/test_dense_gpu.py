"""GPU parity of the dense-basis (non-tensor) operator path -- the one Palace takes for every vector
element and every simplex (/root/reference/palace/fem/libceed/basis.cpp:40-85) -- against the oracle, fed
with the FULL DofToQuad tables in native dof order exactly as the reference hands them to libCEED. Hex
tables are used because the oracle can produce them without MFEM; the kernel itself is element-agnostic.
The tridiagonal curl-oriented restriction of ND tets/prisms (restriction.cpp:301-368) is exercised with
synthetic small-integer transformations."""
import numpy as np
import pytest

from oracle import pyoracle as O
from tests import common

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu
RTOL = 1e-12


def _dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def _rel(a, b):
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


def _dense_op(ctx, prob, kind, blob, curl_orient=None):
    from palace_b200 import capi

    geom = capi.Geom.general(ctx, prob.qdata_ref)
    if kind == O.H1_DIFFUSION:
        sp = prob.h1
        _, grad, _ = O.h1_hex_tables(sp.p, prob.q1d)
        return capi.Op.create_dense(ctx, geom, kind, sp.ndofs, sp.lex_gid.astype(np.int32), None, None, grad, blob), sp
    sp = prob.nd
    interp, curl, _ = O.nd_hex_tables(sp.p, prob.q1d)
    idx, ori = sp.native_restriction()
    return capi.Op.create_dense(ctx, geom, kind, sp.ndofs, idx, ori, interp, curl, blob, curl_orient=curl_orient), sp


@pytest.mark.parametrize("kind", [O.CURLCURL, O.ND_MASS, O.CURLCURL_MASS, O.H1_DIFFUSION])
@pytest.mark.parametrize("p", [1, 2, 3])
def test_dense_apply_and_diagonal_match_oracle(b2p_ctx, p, kind):
    prob = common.make_problem(n=(3, 3, 2), p=p, n_attr=3)  # 18 elements: two full batches of 8 and a ragged one
    blob = common.coefficient(kind, 3, "matrix", a_mass=1.3, a_curl=0.7)
    op, sp = _dense_op(b2p_ctx, prob, kind, blob)
    x = np.random.default_rng(1).random(sp.ndofs)
    y_ref = common.oracle_apply(prob, kind, blob, x)
    yd = torch.full((sp.ndofs,), 3.0, dtype=torch.float64, device="cuda")
    op.apply(_dev(x), yd)
    assert _rel(yd.cpu().numpy(), y_ref) < RTOL
    op.apply_add_ex(-0.5, _dev(x), yd)
    assert _rel(yd.cpu().numpy(), 0.5 * y_ref) < RTOL
    dd = torch.zeros(sp.ndofs, dtype=torch.float64, device="cuda")
    op.diag_add(dd)
    assert _rel(dd.cpu().numpy(), common.oracle_diag(prob, kind, blob)) < RTOL


def test_dense_and_sum_factorised_kernels_agree(b2p_ctx):
    prob = common.make_problem(p=3)
    kind = O.CURLCURL_MASS
    blob = common.coefficient(kind, 3, "matrix")
    dense, sp = _dense_op(b2p_ctx, prob, kind, blob)
    g = common.gpu_geom(b2p_ctx, prob)
    fast = common.gpu_op(b2p_ctx, g, prob, kind, blob)
    x = _dev(np.random.default_rng(2).random(sp.ndofs))
    y1, y2 = torch.empty_like(x), torch.empty_like(x)
    dense.apply(x, y1)
    fast.apply(x, y2)
    assert _rel(y1.cpu().numpy(), y2.cpu().numpy()) < RTOL


def test_curl_oriented_restriction(b2p_ctx):
    """y = sum_e scatter(T_e^T A_e T_e x[idx_e]) with row-major tridiagonal int8 T_e (libCEED
    CeedElemRestrictionCreateCurlOriented semantics as Palace fills it, restriction.cpp:301-329)."""
    prob = common.make_problem(n=(3, 2, 2), p=2, n_attr=2)
    kind = O.CURLCURL_MASS
    blob = common.coefficient(kind, 2, "matrix")
    sp = prob.nd
    P, ne = sp.P, prob.mesh.ne
    rng = np.random.default_rng(3)
    co = np.zeros((ne, P, 3), dtype=np.int8)
    co[:, :, 1] = rng.choice([-1, 1], size=(ne, P))
    # sprinkle 2x2 face-dof style blocks [[a, b], [c, d]] on consecutive pairs
    for e in range(ne):
        for j in rng.choice(np.arange(0, P - 1, 2), size=P // 6, replace=False):
            blk = rng.integers(-1, 2, size=(2, 2))
            co[e, j, 1], co[e, j, 2] = blk[0, 0], blk[0, 1]
            co[e, j + 1, 0], co[e, j + 1, 1] = blk[1, 0], blk[1, 1]
    idx, _ = sp.native_restriction()
    op, _ = _dense_op(b2p_ctx, prob, kind, blob, curl_orient=co)
    interp, curl, _ = O.nd_hex_tables(sp.p, prob.q1d)
    Ae = O.element_matrices(kind, interp, curl, None, prob.qdata_ref, blob, P)
    x = rng.random(sp.ndofs)
    y_ref = np.zeros(sp.ndofs)
    for e in range(ne):
        T = np.zeros((P, P))
        for r in range(P):
            T[r, r] = co[e, r, 1]
            if r > 0:
                T[r, r - 1] = co[e, r, 0]
            if r < P - 1:
                T[r, r + 1] = co[e, r, 2]
        np.add.at(y_ref, idx[e], T.T @ (Ae[e] @ (T @ x[idx[e]])))
    yd = torch.empty(sp.ndofs, dtype=torch.float64, device="cuda")
    op.apply(_dev(x), yd)
    assert _rel(yd.cpu().numpy(), y_ref) < RTOL
