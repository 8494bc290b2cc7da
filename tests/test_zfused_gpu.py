"""GPU parity of the FUSED complex element operator (nd_hex_apply4_kernel<..., CPLX>): a complex sum
sum_i (c_i^r + i c_i^i) A_i of curl-curl / mass operators applied to a split complex vector in ONE pass over the
geometry, against (a) the term-by-term path -- four real applies per term, the reference's
ComplexWrapperOperator, /root/reference/palace/linalg/operator.cpp:98-134 -- and (b) complex arithmetic on
the oracle's assembled matrices. Tolerance 1e-12 relative (FP64, different summation order)."""
import numpy as np
import pytest

from oracle import pyoracle as O
from tests import common

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu
RTOL = 1e-12


def _dev(a):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).cuda()


def _rel(a, b):
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


def _terms(ctx, geom, prob, n_attr, which):
    mk = lambda kind, **kw: (kind, common.coefficient(kind, n_attr, **kw))
    spec = {
        "K+M lossy": [(mk(O.CURLCURL, coeff_type="const"), 1.0 + 0.0j), (mk(O.ND_MASS, coeff_type="matrix"), -9.0 * (1 - 0.05j))],
        "four terms": [(mk(O.CURLCURL, coeff_type="matrix"), 1.0 + 0.0j), (mk(O.ND_MASS, coeff_type="matrix", a_mass=1.7), -2.5 + 0.4j),
                       (mk(O.ND_MASS, coeff_type="scalar", a_mass=0.3), 1.3j), (mk(O.CURLCURL_MASS, coeff_type="matrix"), 0.2 - 0.1j)],
        "real coefficients": [(mk(O.CURLCURL_MASS, coeff_type="matrix"), 0.7 + 0.0j), (mk(O.ND_MASS, coeff_type="const"), 2.0 + 0.0j)],
        "mass only": [(mk(O.ND_MASS, coeff_type="matrix"), 0.5 - 1.5j)],
        "curl only": [(mk(O.CURLCURL, coeff_type="matrix"), -0.5 + 0.25j)],
    }[which]
    ops = [common.gpu_op(ctx, geom, prob, kind, blob) for (kind, blob), _ in spec]
    mats = [common.oracle_matrix(prob, kind, blob, eliminate=False) for (kind, blob), _ in spec]
    return ops, [c for _, c in spec], mats


@pytest.mark.parametrize("which", ["K+M lossy", "four terms", "real coefficients", "mass only", "curl only"])
@pytest.mark.parametrize("p", [1, 2, 3])
def test_fused_complex_apply(b2p_ctx, monkeypatch, p, which):
    from palace_b200 import capi

    n_attr = 3
    prob = common.make_problem(n=(3, 2, 3) if p < 3 else (3, 2, 2), p=p, n_attr=n_attr)  # ragged last batch at every p
    geom = common.gpu_geom(b2p_ctx, prob)
    nd = prob.nd
    ops, coefs, mats = _terms(b2p_ctx, geom, prob, n_attr, which)
    monkeypatch.setenv("B2P_COMPLEX_FUSED", "0")
    A_terms = capi.ComplexOperator.par(b2p_ctx, nd.ndofs, nd.ndofs, ops, coefs, nd.ess_dofs, 1)
    monkeypatch.setenv("B2P_COMPLEX_FUSED", "1")
    A_fused = capi.ComplexOperator.par(b2p_ctx, nd.ndofs, nd.ndofs, ops, coefs, nd.ess_dofs, 1)
    rng = np.random.default_rng(7)
    x = rng.random(nd.ndofs) + 1j * rng.random(nd.ndofs)
    xr, xi = _dev(x.real), _dev(x.imag)
    out = {}
    for name, A in (("terms", A_terms), ("fused", A_fused)):
        yr, yi = torch.full_like(xr, 3.0), torch.full_like(xr, -2.0)
        A.mult(xr, xi, yr, yi)
        hr, hi = torch.empty_like(xr), torch.empty_like(xr)
        A.mult_hermitian_transpose(xr, xi, hr, hi)
        out[name] = (yr.cpu().numpy() + 1j * yi.cpu().numpy(), hr.cpu().numpy() + 1j * hi.cpu().numpy())
    assert A_terms.fused_applies() == 0 and A_fused.fused_applies() == 2   # the fused kernel really ran
    assert _rel(out["fused"][0], out["terms"][0]) < RTOL
    assert _rel(out["fused"][1], out["terms"][1]) < RTOL
    # against complex arithmetic on the oracle's assembled matrices, essential rows DIAG_ONE (rap.cpp:481-517)
    Ao = sum(c * M for c, M in zip(coefs, mats)).tolil()
    ess = nd.ess_dofs
    Ao[ess, :] = 0
    Ao[:, ess] = 0
    Ao[ess, ess] = 1.0
    assert _rel(out["fused"][0], Ao.tocsr() @ x) < RTOL


def test_fused_path_is_skipped_when_terms_differ(b2p_ctx, monkeypatch):
    """Terms on different spaces / orders cannot share one pass: the operator silently keeps the term-by-term path."""
    from palace_b200 import capi

    prob = common.make_problem(n=(2, 2, 2), p=4, n_attr=1)   # q1d = 5: one element slot per warp -> not eligible
    geom = common.gpu_geom(b2p_ctx, prob)
    nd = prob.nd
    blob = common.coefficient(O.CURLCURL_MASS, 1, "const")
    op = common.gpu_op(b2p_ctx, geom, prob, O.CURLCURL_MASS, blob)
    monkeypatch.setenv("B2P_COMPLEX_FUSED", "1")
    A = capi.ComplexOperator.par(b2p_ctx, nd.ndofs, nd.ndofs, [op], [1.0 + 0.5j], nd.ess_dofs, 1)
    x = np.random.default_rng(1).random(nd.ndofs)
    yr, yi = torch.empty(nd.ndofs, dtype=torch.float64, device="cuda"), torch.empty(nd.ndofs, dtype=torch.float64, device="cuda")
    A.mult(_dev(x), _dev(0 * x), yr, yi)
    assert A.fused_applies() == 0
    Ao = common.oracle_matrix(prob, O.CURLCURL_MASS, blob, eliminate=True)
    ref = (1.0 + 0.5j) * (Ao @ x)
    ref[nd.ess_dofs] = x[nd.ess_dofs]
    assert _rel(yr.cpu().numpy() + 1j * yi.cpu().numpy(), ref) < RTOL


@pytest.mark.parametrize("fused", ["0", "1"])
def test_frequency_sweep_updates_coefficients_in_place(b2p_ctx, monkeypatch, fused):
    """A(w) = K - w^2 (1 - i tan d) M + i w C: new coefficients per frequency without rebuilding the operator
    (b2p_coperator_set_coefficients), term-by-term and fused."""
    from palace_b200 import capi

    prob = common.make_problem(n=(3, 2, 2), p=3, n_attr=2)
    geom = common.gpu_geom(b2p_ctx, prob)
    nd = prob.nd
    specs = [(O.CURLCURL, common.coefficient(O.CURLCURL, 2, "const")), (O.ND_MASS, common.coefficient(O.ND_MASS, 2, "matrix")),
             (O.ND_MASS, common.coefficient(O.ND_MASS, 2, "scalar", a_mass=0.3))]
    ops = [common.gpu_op(b2p_ctx, geom, prob, k, b) for k, b in specs]
    mats = [common.oracle_matrix(prob, k, b, eliminate=False) for k, b in specs]
    coefs_of = lambda w: [1.0 + 0.0j, -w * w * (1 - 0.02j), 1j * w]
    monkeypatch.setenv("B2P_COMPLEX_FUSED", fused)
    A = capi.ComplexOperator.par(b2p_ctx, nd.ndofs, nd.ndofs, ops, coefs_of(1.0), nd.ess_dofs, 1)
    rng = np.random.default_rng(3)
    x = rng.random(nd.ndofs) + 1j * rng.random(nd.ndofs)
    xr, xi = _dev(x.real), _dev(x.imag)
    for w in (1.0, 2.5, 0.4):
        A.set_coefficients(coefs_of(w))
        yr, yi = torch.empty_like(xr), torch.empty_like(xr)
        A.mult(xr, xi, yr, yi)
        Ao = sum(c * M for c, M in zip(coefs_of(w), mats)).tolil()
        ess = nd.ess_dofs
        Ao[ess, :] = 0
        Ao[:, ess] = 0
        Ao[ess, ess] = 1.0
        assert _rel(yr.cpu().numpy() + 1j * yi.cpu().numpy(), Ao.tocsr() @ x) < RTOL
    assert A.fused_applies() == (3 if fused == "1" else 0)


@pytest.mark.parametrize("kind", [O.CURLCURL, O.ND_MASS, O.CURLCURL_MASS])
@pytest.mark.parametrize("p", [1, 3, 4])
def test_two_vector_apply(b2p_ctx, p, kind):
    """Two right-hand sides in one pass over the geometry (b2p_op_apply_add_pair; p = 4 takes the two-apply fallback)."""
    prob = common.make_problem(n=(3, 2, 2), p=p, n_attr=3)
    geom = common.gpu_geom(b2p_ctx, prob)
    blob = common.coefficient(kind, 3, "matrix")
    op = common.gpu_op(b2p_ctx, geom, prob, kind, blob)
    op.set_essential(prob.nd.ess_dofs.astype(np.int32))
    rng = np.random.default_rng(2)
    x0, x1 = rng.random(prob.nd.ndofs), rng.random(prob.nd.ndofs)
    for masked in (False, True):
        y0, y1 = torch.full((prob.nd.ndofs,), 1.0, dtype=torch.float64, device="cuda"), torch.full((prob.nd.ndofs,), -2.0, dtype=torch.float64, device="cuda")
        op.apply_add_pair(0.5, _dev(x0), _dev(x1), y0, y1, masked=masked)
        r0, r1 = torch.full_like(y0, 1.0), torch.full_like(y1, -2.0)
        op.apply_add_ex(0.5, _dev(x0), r0, masked=masked)
        op.apply_add_ex(0.5, _dev(x1), r1, masked=masked)
        assert _rel(y0.cpu().numpy(), r0.cpu().numpy()) < RTOL and _rel(y1.cpu().numpy(), r1.cpu().numpy()) < RTOL
    assert _rel(r0.cpu().numpy() - 1.0, 0.5 * common.oracle_matrix(prob, kind, blob, eliminate=False).tolil()[:, :].tocsr().dot(
        np.where(np.isin(np.arange(prob.nd.ndofs), prob.nd.ess_dofs), 0.0, x0)) * ~np.isin(np.arange(prob.nd.ndofs), prob.nd.ess_dofs)) < RTOL
