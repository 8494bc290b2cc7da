"""GPU: malformed input to the entry points added at the end of round 2 comes back as an error code with a message
(b2p_last_error), never as a crash -- the contract of include/b2p.h (PalaceCeedCall-style, ceed.hpp:13-33)."""
import numpy as np
import pytest
import scipy.sparse as sparse

from oracle import pyoracle as O
from tests import common

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def parts(b2p_ctx):
    from palace_b200 import capi

    prob = common.make_problem(n=(2, 2, 2), p=1, n_attr=1)
    geom = capi.Geom.general(b2p_ctx, prob.qdata_ref)
    sp = prob.nd
    interp, curl, _ = O.nd_hex_tables(sp.p, prob.q1d)
    idx, ori = sp.native_restriction()
    return prob, geom, sp, interp, curl, idx, ori


def test_mixed_curl_kinds_need_both_tables_and_a_general_geometry(b2p_ctx, parts):
    from palace_b200 import capi

    prob, geom, sp, interp, curl, idx, ori = parts
    blob = common.coefficient(O.ND_MASS, 1, "const")
    with pytest.raises(capi.B2PError, match="missing tables"):
        capi.Op.create_dense(b2p_ctx, geom, capi.ND_WEAKCURL, sp.ndofs, idx, ori, interp, None, blob)
    with pytest.raises(capi.B2PError, match="bad kind"):
        capi.Op.create_dense(b2p_ctx, geom, 6, sp.ndofs, idx, ori, interp, curl, blob)
    hexgeom = common.gpu_geom(b2p_ctx, prob)
    with pytest.raises(capi.B2PError, match="general geometry"):
        capi.Op.create_dense(b2p_ctx, hexgeom, capi.ND_MIXEDCURL, sp.ndofs, idx, ori, interp, curl, blob)


def test_sparse_matrix_and_triple_products_check_their_shapes(b2p_ctx, parts):
    import ctypes as C

    from palace_b200 import capi

    prob, geom, sp, interp, curl, idx, ori = parts
    n = sp.ndofs
    L = capi.lib()
    # malformed CSR arrays
    h = C.c_void_p()
    rp = np.array([0, 2, 1], dtype=np.int32)
    cl = np.array([0, 1], dtype=np.int32)
    vl = np.array([1.0, 1.0])
    assert L.b2p_spmat_create(b2p_ctx.h, C.c_int64(2), C.c_int64(2), rp.ctypes.data_as(C.c_void_p), cl.ctypes.data_as(C.c_void_p),
                              vl.ctypes.data_as(C.c_void_p), C.byref(h)) == 1
    rp = np.array([0, 1, 2], dtype=np.int32)
    cl = np.array([0, 5], dtype=np.int32)
    assert L.b2p_spmat_create(b2p_ctx.h, C.c_int64(2), C.c_int64(2), rp.ctypes.data_as(C.c_void_p), cl.ctypes.data_as(C.c_void_p),
                              vl.ctypes.data_as(C.c_void_p), C.byref(h)) == 1
    # products that do not chain
    blob = common.coefficient(O.ND_MASS, 1, "const")
    M = capi.Op.create_dense(b2p_ctx, geom, capi.ND_MASS, n, idx, ori, interp, None, blob)
    A = capi.Operator.par(b2p_ctx, n, n, [M], None, None, diag_policy=1)
    Pbad = capi.SpMat(b2p_ctx, sparse.identity(n + 1, format="csr"))
    with pytest.raises(capi.B2PError, match="prolongation has"):
        capi.operator_rap(b2p_ctx, A, Pbad)
    with pytest.raises(capi.B2PError, match="do not chain"):
        capi.operator_triple(b2p_ctx, Pbad, A, None)
    Pok = capi.SpMat(b2p_ctx, sparse.identity(n, format="csr"))
    with pytest.raises(capi.B2PError, match="essential true dof"):
        capi.operator_rap(b2p_ctx, A, Pok, ess_tdofs=np.array([n + 3]))
    # an empty matrix is legal
    E = capi.SpMat(b2p_ctx, sparse.csr_matrix((3, 4)))
    y = torch.full((3,), 7.0, dtype=torch.float64, device="cuda")
    E.mult(torch.ones(4, dtype=torch.float64, device="cuda"), y)
    assert float(y.abs().max()) == 0.0
