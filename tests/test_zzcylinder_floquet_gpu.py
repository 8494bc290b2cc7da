"""GPU, Floquet-periodic terms against the reference's own output: the device dense-basis operators of the reference's
examples/cylinder/floquet.json -- curl-curl, the [k x]^T [k x] mass, MixedVectorWeakCurl and MixedVectorCurl with mu^-1 [k x]
(B2P_ND_WEAKCURL / B2P_ND_MIXEDCURL), the mass -- on its quadratic tetrahedral mesh at Nedelec order 4
(a) equal the oracle's assembled matrices on random vectors to 1e-12 and (b) combined into ONE complex operator
K_r + i K_i - lambda M have the REFERENCE's stored eigenfrequencies (test/data/regression/ref/cylinder/floquet/eig.csv) as
eigenvalues: || (K_r + i K_i) v - lambda_ref M v || <= 1e-6 || K_r v || for the eigenvectors of the oracle-side pencil
(tests/test_cylinder_floquet_golden.py holds that pencil to the stored numbers at 1e-11 ... 5e-8)."""
import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from oracle import pyoracle as O
from palace_b200.host import coeff as cf
from tests import common
from tests.test_cylinder_floquet_golden import assembled, cross_matrix, floquet_matrices
from tests.test_cylinder_tet_golden import C0, FIX, sigma_target, space_and_tables

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def _dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def _rel(a, b):
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


def test_device_floquet_operators_have_the_reference_eigenpairs(b2p_ctx):
    from palace_b200 import capi

    p = int(FIX["order"])
    mesh, nd, interp, curl, qd = space_and_tables(p)
    k = FIX["floquet_wave_vector"]
    kx = cross_matrix(k)
    am = np.array([0])
    one = cf.coeff_ctx(a=1.0)
    blobs = {"K": (capi.CURLCURL, one), "Mk": (capi.ND_MASS, cf.coeff_ctx(am, (kx.T @ kx)[None])),
             "W": (capi.ND_WEAKCURL, cf.coeff_ctx(am, kx[None], a=-1.0)), "C": (capi.ND_MIXEDCURL, cf.coeff_ctx(am, kx[None], transpose=True)),
             "M": (capi.ND_MASS, one)}
    geom = capi.Geom.general(b2p_ctx, qd)
    n = nd.ndofs
    ops = {}
    for name, (kind, blob) in blobs.items():
        ops[name] = capi.Op.create_dense(b2p_ctx, geom, kind, n, nd.idx, None, interp if kind != capi.CURLCURL else None,
                                         curl if kind != capi.ND_MASS else None, blob, curl_orient=nd.curl_orient)
    x = np.random.default_rng(3).standard_normal(n)
    y = torch.empty(n, dtype=torch.float64, device="cuda")
    for name, (kind, blob) in blobs.items():
        ops[name].apply(_dev(x), y)
        assert _rel(y.cpu().numpy(), assembled(nd, kind, interp, curl, qd, blob) @ x) < 1e-12, name
    # eigenvectors of the oracle-side Hermitian pencil (real doubled form), then the device complex operator on them
    Kr, Ki, M = floquet_matrices(nd, interp, curl, qd, k)
    free = np.setdiff1d(np.arange(n), nd.ess_dofs)
    Krf, Kif, Mf = Kr[free][:, free], Ki[free][:, free], M[free][:, free]
    lam2, V2 = common.eigsh_above(sp.bmat([[Krf, -Kif], [Kif, Krf]], format="csc"), sp.block_diag([Mf, Mf], format="csc"), 30,
                                  sigma_target(), extra=8, vectors=True)
    order = np.argsort(lam2)
    f_ref = FIX["floquet_f_re_ghz"] + 1j * FIX["floquet_f_im_ghz"]
    lam_ref = ((2 * np.pi * f_ref * 1e9 * float(FIX["L0"]) / C0) ** 2 * float(FIX["eps_r"]) * (1 - 1j * float(FIX["loss_tan"]))).real
    Z = capi.ComplexOperator.par(b2p_ctx, n, n, [ops["K"], ops["Mk"], ops["W"], ops["C"], ops["M"]], [1.0, 1.0, 1.0j, 1.0j, 0.0],
                                 nd.ess_dofs, diag_policy=0)
    KR = capi.ComplexOperator.par(b2p_ctx, n, n, [ops["K"], ops["Mk"]], [1.0, 1.0], nd.ess_dofs, diag_policy=0)
    yr, yi, kr, ki_ = (torch.empty(n, dtype=torch.float64, device="cuda") for _ in range(4))
    worst = 0.0
    for j in range(15):
        v2 = V2[:, order[2 * j]]
        v = np.zeros(n, dtype=complex)
        v[free] = v2[:free.size] + 1j * v2[free.size:]
        Z.set_coefficients([1.0, 1.0, 1.0j, 1.0j, -lam_ref[j]])
        Z.mult(_dev(v.real), _dev(v.imag), yr, yi)
        KR.mult(_dev(v.real), _dev(v.imag), kr, ki_)
        r = np.linalg.norm((yr.cpu().numpy() + 1j * yi.cpu().numpy())[free])
        worst = max(worst, r / np.linalg.norm((kr.cpu().numpy() + 1j * ki_.cpu().numpy())[free]))
    print("max || (K_r + i K_i - lambda_ref M) v || / || K_r v || over the 15 reference modes:", worst)
    assert worst < 1e-6
