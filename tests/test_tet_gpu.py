"""GPU parity of the Nedelec TETRAHEDRON path (BASELINE configs 3 and 4 are tet meshes): the dense-basis
DMMA operator (b2p_op_create_dense) fed with real simplex tables, the tridiagonal curl-oriented restriction
of scrambled tets (/root/reference/palace/fem/libceed/restriction.cpp:301-368) and prebuilt q-data of straight
and curved tets, against the oracle. Bit-level agreement is not expected (different summation order):
||dy|| <= 1e-12 ||y||, the reference's own bound (test/unit/test-libceed.cpp:262-268)."""
import numpy as np
import pytest

from oracle import pyoracle as O
from palace_b200.host import coeff as cf
from palace_b200.host import tetspace as ts

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu
RTOL = 1e-12


def _dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def _rel(a, b):
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


def _problem(p, n, geom_order=1, warp=0.0, n_attr=3, degree=None, seed=5):
    mesh = ts.box_tet_mesh(n, (1.0, 0.8, 0.9), jitter=0.25, scramble_seed=seed, n_attr=n_attr, warp_amp=warp)
    sp = ts.build_nd_tet_space(mesh, p)
    interp, curl, qpts, qw = ts.nd_tet_tables(p, degree)
    qd = ts.geom_qdata(mesh.node_coords(geom_order), mesh.attr, geom_order, qpts, qw)
    return mesh, sp, interp, curl, qd


def _blob(kind, n_attr):
    am, mc = cf.test_suite_coefficient(n_attr, "matrix")
    first = cf.coeff_ctx(am, mc, a=1.3)
    second = cf.coeff_ctx(am, mc[::-1].copy() if len(mc) > 1 else mc, a=0.7, transpose=True)
    return cf.coeff_ctx_pair(first, second) if kind == O.CURLCURL_MASS else (second if kind == O.CURLCURL else first)


def _op(ctx, kind, sp, interp, curl, qd, blob, signs_only=False):
    from palace_b200 import capi

    geom = capi.Geom.general(ctx, qd)
    if signs_only:
        return capi.Op.create_dense(ctx, geom, kind, sp.ndofs, sp.idx, sp.orient_signs(), interp, curl, blob)
    return capi.Op.create_dense(ctx, geom, kind, sp.ndofs, sp.idx, None, interp, curl, blob, curl_orient=sp.curl_orient)


@pytest.mark.parametrize("kind", [O.CURLCURL, O.ND_MASS, O.CURLCURL_MASS])
@pytest.mark.parametrize("p", [1, 2, 3])
def test_tet_apply_matches_oracle(b2p_ctx, p, kind):
    mesh, sp, interp, curl, qd = _problem(p, (2, 2, 1))       # 24 tets: three full batches of 8
    blob = _blob(kind, 3)
    op = _op(b2p_ctx, kind, sp, interp, curl, qd, blob)
    x = np.random.default_rng(1).random(sp.ndofs)
    y_ref = O.apply_add_co(kind, interp, curl, sp.idx, sp.curl_orient, qd, blob, x, np.zeros(sp.ndofs))
    yd = torch.full((sp.ndofs,), 7.0, dtype=torch.float64, device="cuda")
    op.apply(_dev(x), yd)                                      # Mult: zero-fill, then add
    assert _rel(yd.cpu().numpy(), y_ref) < RTOL
    op.apply_add_ex(-0.25, _dev(x), yd)                        # AddMult with a coefficient
    assert _rel(yd.cpu().numpy(), 0.75 * y_ref) < RTOL


def test_tet_p1_sign_orientation_equals_curl_orientation(b2p_ctx):
    mesh, sp, interp, curl, qd = _problem(1, (2, 1, 1))       # 12 tets: ragged last batch
    kind, blob = O.CURLCURL_MASS, _blob(O.CURLCURL_MASS, 3)
    a = _op(b2p_ctx, kind, sp, interp, curl, qd, blob, signs_only=True)
    b = _op(b2p_ctx, kind, sp, interp, curl, qd, blob)
    x = _dev(np.random.default_rng(2).random(sp.ndofs))
    ya, yb = torch.empty_like(x), torch.empty_like(x)
    a.apply(x, ya)
    b.apply(x, yb)
    assert _rel(ya.cpu().numpy(), yb.cpu().numpy()) < RTOL
    y_ref = O.apply_add(kind, interp, curl, sp.idx, sp.orient_signs(), qd, blob, x.cpu().numpy(), np.zeros(sp.ndofs))
    assert _rel(ya.cpu().numpy(), y_ref) < RTOL


def test_curved_tets_p3(b2p_ctx):
    """Quadratic (curved) tets, over-integrated: non-constant Jacobians through the prebuilt q-data path."""
    mesh, sp, interp, curl, qd = _problem(3, (1, 2, 1), geom_order=2, warp=0.03, degree=8)
    kind, blob = O.CURLCURL_MASS, _blob(O.CURLCURL_MASS, 3)
    op = _op(b2p_ctx, kind, sp, interp, curl, qd, blob)
    x = np.random.default_rng(3).random(sp.ndofs)
    y_ref = O.apply_add_co(kind, interp, curl, sp.idx, sp.curl_orient, qd, blob, x, np.zeros(sp.ndofs))
    yd = torch.empty(sp.ndofs, dtype=torch.float64, device="cuda")
    op.apply(_dev(x), yd)
    assert _rel(yd.cpu().numpy(), y_ref) < RTOL


def test_tet_p6_apply_and_symmetry(b2p_ctx):
    """BASELINE config 4's order: P = 216 dofs per tet, 343-point rule (largest element the dense kernel stages)."""
    mesh, sp, interp, curl, qd = _problem(6, (1, 1, 1), n_attr=1)
    kind = O.CURLCURL_MASS
    blob = cf.coeff_ctx_pair(cf.coeff_ctx(a=1.0), cf.coeff_ctx(a=1.0))
    op = _op(b2p_ctx, kind, sp, interp, curl, qd, blob)
    rng = np.random.default_rng(4)
    x, z = rng.random(sp.ndofs), rng.random(sp.ndofs)
    y_ref = O.apply_add_co(kind, interp, curl, sp.idx, sp.curl_orient, qd, blob, x, np.zeros(sp.ndofs))
    yx, yz = torch.empty(sp.ndofs, dtype=torch.float64, device="cuda"), torch.empty(sp.ndofs, dtype=torch.float64, device="cuda")
    op.apply(_dev(x), yx)
    op.apply(_dev(z), yz)
    assert _rel(yx.cpu().numpy(), y_ref) < 1e-11
    assert abs(z @ yx.cpu().numpy() - x @ yz.cpu().numpy()) < 1e-11 * abs(z @ yx.cpu().numpy())


def test_tet_essential_rows_and_pec_box_energy(b2p_ctx):
    """ParOperator over the tet operator with the PEC boundary eliminated (rap.cpp:207-233 semantics): rows of
    essential dofs return x, interior rows ignore essential inputs."""
    from palace_b200 import capi

    mesh, sp, interp, curl, qd = _problem(2, (2, 2, 2), n_attr=1)
    kind = O.CURLCURL_MASS
    blob = cf.coeff_ctx_pair(cf.coeff_ctx(a=1.0), cf.coeff_ctx(a=1.0))
    op = _op(b2p_ctx, kind, sp, interp, curl, qd, blob)
    A = capi.Operator.par(b2p_ctx, sp.ndofs, sp.ndofs, [op], None, sp.ess_dofs, diag_policy=1)
    x = np.random.default_rng(5).random(sp.ndofs)
    xm = x.copy()
    xm[sp.ess_dofs] = 0.0
    y_ref = O.apply_add_co(kind, interp, curl, sp.idx, sp.curl_orient, qd, blob, xm, np.zeros(sp.ndofs))
    y_ref[sp.ess_dofs] = x[sp.ess_dofs]
    yd = torch.empty(sp.ndofs, dtype=torch.float64, device="cuda")
    A.mult(_dev(x), yd)
    assert _rel(yd.cpu().numpy(), y_ref) < RTOL
