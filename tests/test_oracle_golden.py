"""CPU: pins the oracle's restated pointwise arithmetic (geometry factors, D for hcurl / hdiv /
hdivmass) to golden vectors produced by the REFERENCE's own QFunction headers
(tests/golden/make_golden.py; /root/reference/palace/fem/qfunctions/33/*.h), and -- when oracle/_ref
is present -- to the compiled reference directly."""
import ctypes as C
import os

import numpy as np
import pytest

from oracle import pyoracle as O

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "qf_golden.npz"))
TOL = 2e-15


def _rel(a, b):
    return np.abs(a - b).max() / np.abs(b).max()


def test_geom_factor_matches_reference_golden():
    # oracle geometry from J: restated in orc_geom_hex_qdata; here re-derive q-data from the golden J
    J, qw, attr = G["J"], G["qw"], G["attr"]
    Q = J.shape[1]
    qd = np.empty((11, Q))
    for i in range(Q):
        M = J[:, i].reshape(3, 3, order="F")
        det = np.linalg.det(M)
        qd[0, i] = attr[i]
        qd[1, i] = qw[i] * det
        qd[2:, i] = np.linalg.inv(M).T.ravel(order="F")  # adj(J)^T/detJ = J^-T
    assert _rel(qd, G["qdata"]) < 1e-13


@pytest.mark.parametrize("case", ["hcurl", "hcurl_identity", "hdiv", "hdivmass"])
def test_pointwise_D_matches_reference_golden(case):
    qdata, u, c = np.ascontiguousarray(G["qdata"]), np.ascontiguousarray(G["u"]), np.ascontiguousarray(G["c"])
    if case == "hcurl":
        v, _ = O.apply_D(O.ND_MASS, np.ascontiguousarray(G["ctx_mass"]), qdata, u, None)
        assert _rel(v, G["hcurl_v"]) < TOL
    elif case == "hcurl_identity":
        v, _ = O.apply_D(O.ND_MASS, np.ascontiguousarray(G["ctx_id"]), qdata, u, None)
        assert _rel(v, G["hcurl_v_identity"]) < TOL
        # H1 diffusion uses the same QFunction on the gradient (integ/diffusion.cpp:37-42)
        _, w = O.apply_D(O.H1_DIFFUSION, np.ascontiguousarray(G["ctx_id"]), qdata, None, u)
        assert _rel(w, G["hcurl_v_identity"]) < TOL
    elif case == "hdiv":
        _, w = O.apply_D(O.CURLCURL, np.ascontiguousarray(G["ctx_curl"]), qdata, None, c)
        assert _rel(w, G["hdiv_w"]) < TOL
    else:
        v, w = O.apply_D(O.CURLCURL_MASS, np.ascontiguousarray(G["ctx_pair"]), qdata, u, c)
        assert _rel(v, G["hdivmass_v"]) < TOL and _rel(w, G["hdivmass_w"]) < TOL


def test_build_qfunctions_are_consistent_with_apply():
    """The reference's assembled-q-data QFunctions (f_build_*_33 + f_apply_33) give the same v as
    the on-the-fly ones: checks the golden build outputs against the golden apply outputs."""
    u, c = G["u"], G["c"]
    qd = G["build_hdivmass"]
    v = np.einsum("rcq,cq->rq", qd[:9].reshape(3, 3, -1, order="F"), u)
    w = np.einsum("rcq,cq->rq", qd[9:].reshape(3, 3, -1, order="F"), c)
    assert _rel(v, G["hdivmass_v"]) < 1e-14 and _rel(w, G["hdivmass_w"]) < 1e-14
    assert _rel(np.einsum("rcq,cq->rq", G["build_hcurl"].reshape(3, 3, -1, order="F"), u), G["hcurl_v"]) < 1e-14
    assert _rel(np.einsum("rcq,cq->rq", G["build_hdiv"].reshape(3, 3, -1, order="F"), c), G["hdiv_w"]) < 1e-14


@pytest.mark.skipif(O.ref() is None, reason="oracle/_ref not built (no /root/reference on this box)")
def test_against_compiled_reference_headers():
    ref = O.ref()
    rng = np.random.default_rng(7)
    Q = 40
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    qdata = np.ascontiguousarray(G["qdata"][:, :Q])
    u, c = rng.random((3, Q)), rng.random((3, Q))
    v, w = np.empty((3, Q)), np.empty((3, Q))
    ctx = np.ascontiguousarray(G["ctx_pair"])
    assert ref.ref_apply_hdivmass_33(p(ctx), Q, p(qdata), p(u), p(c), p(v), p(w)) == 0
    vo, wo = O.apply_D(O.CURLCURL_MASS, ctx, qdata, u, c)
    assert _rel(vo, v) < TOL and _rel(wo, w) < TOL
