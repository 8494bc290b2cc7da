"""GPU, flux error estimators against the reference's own output: the device TimeDependentFluxErrorEstimator pipeline
(b2p_flux_estimator_* in the grad-flux and curl-flux configurations + b2p_flux_estimator_sqrt_scale, b2p_operator_vecfe_mass as the
RT mass) run over the 15 cavity modes of the reference's examples/cylinder/cavity_pec.json reproduces (a) the oracle pipeline of
tests/test_cylinder_indicator_golden.py element by element and (b) at order 4 the global norm the reference's regression suite
stores in error-indicators.csv (1.7317852e-3; the device pipeline gives 1.7317850e-3 on the emulation build)."""
import os

import numpy as np
import pytest

from oracle import pyoracle as O
from palace_b200.host import coeff as cf
from palace_b200.host import hexspace as hs
from tests.test_cylinder_golden import FIX
from tests.test_cylinder_indicator_golden import REF_NORM, cavity_modes, curl_dofs, oracle_indicators

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def _dev(a):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).cuda()


def test_device_estimators_reproduce_the_reference_indicator_norm(b2p_ctx):
    from palace_b200 import capi

    p = int(os.environ.get("B2P_TEST_ORDER", int(FIX["order"])))  # (B2P_TEST_ORDER=2: a quick run, comparison (a) only)
    c = cavity_modes(p)
    mesh, nd, rt, q1d, eps = c["mesh"], c["nd"], c["rt"], c["q1d"], c["eps"]
    ne = mesh.ne
    geom = capi.Geom.from_qdata(b2p_ctx, c["qd"], q1d)
    t = hs.tables_1d(p, q1d)
    mop = capi.Op.create(b2p_ctx, geom, O.ND_MASS, p, nd.ndofs, c["idx"], c["ori"], nd.dof_map, t.Bo, t.Bc, t.Gc, cf.coeff_ctx(a=1.0), False)
    Mnd = capi.Operator.par(b2p_ctx, nd.ndofs, nd.ndofs, [mop], None, None, diag_policy=1)
    sp_rt = dict(P=rt.P, map_type=capi.MAP_HDIV, interp=hs.rt_hex_dense_interp(p, q1d), idx=rt.lex_gid, orient=rt.lex_sign, lsize=rt.ndofs)
    sp_nd = dict(P=nd.P, map_type=capi.MAP_HCURL, interp=hs.nd_hex_dense_interp(p, q1d), idx=nd.lex_gid, orient=nd.lex_sign, lsize=nd.ndofs)
    Mrt = capi.vecfe_mass_operator(b2p_ctx, geom, sp_rt)
    I9 = np.eye(3).ravel()[None]
    curl_est = capi.FluxEstimator(b2p_ctx, geom, sp_rt, sp_nd, I9, I9, I9, Mnd, tol=1e-12, max_it=20000)
    grad_est = capi.FluxEstimator(b2p_ctx, geom, sp_nd, sp_rt, eps * I9, np.sqrt(eps) * I9, I9 / np.sqrt(eps), Mrt, tol=1e-12, max_it=20000)
    acc = np.zeros(ne)
    yr = torch.empty(rt.ndofs, dtype=torch.float64, device="cuda")
    for v, lam in zip(c["modes"], c["lam"]):
        B = curl_dofs(nd, rt, v) / np.sqrt(lam / eps)
        Mnd.mult(_dev(v), (mv := torch.empty(nd.ndofs, dtype=torch.float64, device="cuda")))
        Mrt.mult(_dev(B), yr)
        Et = 0.5 * eps * float(v @ mv.cpu().numpy()) + 0.5 * float(B @ yr.cpu().numpy())
        D = torch.zeros(rt.ndofs, dtype=torch.float64, device="cuda")
        H = torch.zeros(nd.ndofs, dtype=torch.float64, device="cuda")
        est = torch.zeros(ne, dtype=torch.float64, device="cuda")
        grad_est.project(_dev(v), D)
        grad_est.integrate(_dev(v), D, est)
        curl_est.project(_dev(B), H)
        curl_est.integrate(_dev(B), H, est)
        assert grad_est.stats()["converged"] and curl_est.stats()["converged"]
        capi.flux_sqrt_scale(b2p_ctx, ne, 0.5 / Et, est)
        acc += est.cpu().numpy() ** 2
    e = np.sqrt(acc / len(c["modes"]))
    e_ref = oracle_indicators(c)
    assert np.abs(e - e_ref).max() < 1e-7 * e_ref.max()
    print("device indicator norm:", np.linalg.norm(e))
    if p == int(FIX["order"]):
        assert abs(np.linalg.norm(e) / REF_NORM - 1) < 1e-6
