"""GPU, tetrahedral path against the reference's own output: the device dense-basis (DMMA) operators K and M of the Nedelec
order-4 space on the reference's quadratic tet mesh (periodic waveguide, tests/golden/cylinder_waveguide_tet.npz) (a) equal
the oracle's assembled matrices on random and eigen vectors to 1e-12, and (b) have the REFERENCE's stored eigenfrequencies
(test/data/regression/ref/cylinder/waveguide/eig.csv) as eigenvalues: || K v - lambda_ref M v || <= 1e-7 || K v || for the 15
eigenvectors of the pencil -- the residual a wrong dof orientation, table entry or geometry factor on any of the 288 elements
would blow up by orders of magnitude."""
import numpy as np
import pytest
import scipy.sparse.linalg as spla

from oracle import pyoracle as O
from palace_b200.host import coeff as cf
from tests import common
from tests.test_cylinder_tet_golden import C0, FIX, oracle_matrix, sigma_target, space_and_tables

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def _dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def _rel(a, b):
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


def test_device_tet_operators_have_the_reference_eigenpairs(b2p_ctx):
    from palace_b200 import capi

    p = int(FIX["order"])
    mesh, nd, interp, curl, qd = space_and_tables(p)
    one = cf.coeff_ctx(a=1.0)
    geom = capi.Geom.general(b2p_ctx, qd)
    Kd = capi.Op.create_dense(b2p_ctx, geom, capi.CURLCURL, nd.ndofs, nd.idx, None, interp, curl, one, curl_orient=nd.curl_orient)
    Md = capi.Op.create_dense(b2p_ctx, geom, capi.ND_MASS, nd.ndofs, nd.idx, None, interp, curl, one, curl_orient=nd.curl_orient)
    K, M = oracle_matrix(nd, O.CURLCURL, interp, curl, qd), oracle_matrix(nd, O.ND_MASS, interp, curl, qd)
    x = np.random.default_rng(3).standard_normal(nd.ndofs)
    y = torch.empty(nd.ndofs, dtype=torch.float64, device="cuda")
    for A, Ad in ((K, Kd), (M, Md)):
        Ad.apply(_dev(x), y)
        assert _rel(y.cpu().numpy(), A @ x) < 1e-12
    free = np.setdiff1d(np.arange(nd.ndofs), nd.ess_dofs)
    lam, V = common.eigsh_above(K[free][:, free].tocsc(), M[free][:, free].tocsc(), 15, sigma_target(), vectors=True)
    order = np.argsort(lam)
    # the reference's complex frequencies -> eigenvalue of the real pencil: lambda = (2 pi f L0 / c0)^2 eps_r (1 - i tan d)
    f_ref = FIX["ref_f_re_ghz"] + 1j * FIX["ref_f_im_ghz"]
    lam_ref = ((2 * np.pi * f_ref * 1e9 * float(FIX["L0"]) / C0) ** 2 * float(FIX["eps_r"]) * (1 - 1j * float(FIX["loss_tan"]))).real
    ky, my = torch.empty_like(y), torch.empty_like(y)
    worst = 0.0
    for j, i in enumerate(order):
        v = np.zeros(nd.ndofs)
        v[free] = V[:, i]
        Kd.apply(_dev(v), ky)
        Md.apply(_dev(v), my)
        kv, mv = ky.cpu().numpy()[free], my.cpu().numpy()[free]
        assert _rel(kv, (K @ v)[free]) < 1e-11
        worst = max(worst, np.linalg.norm(kv - lam_ref[j] * mv) / np.linalg.norm(kv))
    print("max || K v - lambda_ref M v || / || K v || over the 15 reference modes:", worst)
    assert worst < 1e-7
