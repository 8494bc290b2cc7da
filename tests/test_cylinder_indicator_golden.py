"""End-to-end pin of the FLUX ERROR ESTIMATORS against the reference's own output (CPU, oracle side): besides eig.csv the reference's
regression suite stores error-indicators.csv for examples/cylinder/cavity_pec.json -- the statistics of the element indicators
e_K = sqrt(1/N sum_n eta_Kn^2) over the N = 15 modes (fem/errorindicator.cpp:11-47), each mode contributing
eta_K^2 = (eta_K^2[eps E -> RT] + eta_K^2[mu^-1 B -> ND]) * 0.5 / E_total (TimeDependentFluxErrorEstimator::AddErrorIndicator,
linalg/errorestimator.cpp:531-545, driven from drivers/eigensolver.cpp:440-500 with B = -1/(i omega) curl E).
oracle/estimator.py (mixed mass matrices, projections, element integrals; pointwise arithmetic pinned to the reference's headers)
on the oracle-side hexahedral discretisation of test_cylinder_golden.py reproduces the stored GLOBAL NORM to 1e-7 -- the reference
projects with tolerance 1e-6 -- i.e. both estimators, the discrete curl, the RT space and the energy normalisation as a whole.
The per-element distribution (minimum, maximum, mean) is only reproduced to a few percent: within a degenerate pair the
reference's eigensolver returns some basis of the eigenspace, every vector of which has the same global error (the cylinder's
rotational symmetry) but puts it on different elements."""
import numpy as np
import scipy.sparse.linalg as spla

from oracle import estimator as E
from oracle import pyoracle as O
from oracle import solvers as S
from palace_b200.host import coeff as cf
from palace_b200.host import hexspace as hs
from tests import common
from tests.test_cylinder_golden import FIX, cylinder_problem, target_lambda

# test/data/regression/ref/cylinder/cavity_pec/error-indicators.csv: Norm, Minimum, Maximum, Mean
REF_NORM, REF_MIN, REF_MAX, REF_MEAN = 1.731785223943e-03, 7.014620941947e-05, 3.001205345978e-04, 1.662997622515e-04


def cavity_modes(p, n_modes=15):
    """Discretisation, the n_modes eigenpairs above the target, and everything the estimators need (oracle side)."""
    mesh, topo, nd, q1d, qd = cylinder_problem(p)
    rt = hs.build_rt_space(mesh, topo, p)
    interp, curl, _ = O.nd_hex_tables(p, q1d)
    idx, ori = nd.native_restriction()
    one = cf.coeff_ctx(a=1.0)
    K = S.assemble_sparse(O.element_matrices(O.CURLCURL, interp, curl, ori, qd, one, nd.P), idx.astype(np.int64), nd.ndofs).tocsr()
    M = S.assemble_sparse(O.element_matrices(O.ND_MASS, interp, curl, ori, qd, one, nd.P), idx.astype(np.int64), nd.ndofs).tocsr()
    free = np.setdiff1d(np.arange(nd.ndofs), nd.ess_dofs)
    lam, V = common.eigsh_above(K[free][:, free].tocsc(), M[free][:, free].tocsc(), n_modes, target_lambda(), vectors=True)
    modes = np.zeros((n_modes, nd.ndofs))
    modes[:, free] = V.T
    return dict(mesh=mesh, topo=topo, nd=nd, rt=rt, q1d=q1d, qd=qd, interp=interp, idx=idx, ori=ori, M=M, lam=lam, modes=modes,
                eps=float(FIX["eps_r"]))


def curl_dofs(nd, rt, x):
    """B = Curl E on the RT dofs (the discrete curl of the host layer)."""
    Cl = hs.discrete_curl_matrix(nd.p)
    B = np.zeros(rt.ndofs)
    for e in range(nd.lex_gid.shape[0]):
        B[rt.lex_gid[e]] = rt.lex_sign[e] * (Cl @ (nd.lex_sign[e] * x[nd.lex_gid[e]]))
    return B


def oracle_indicators(c):
    """e_K over the modes of c (dict of cavity_modes), all on assembled matrices with direct solves."""
    mesh, nd, rt, qd, interp, eps = c["mesh"], c["nd"], c["rt"], c["qd"], c["interp"], c["eps"]
    ne = mesh.ne
    I3 = [np.eye(3)] * ne
    rt_interp = E.rt_hex_tables(nd.p, c["q1d"])
    idx_n, ori_n = c["idx"].astype(np.int64), c["ori"].astype(float)
    idx_r, ori_r = rt.lex_gid, rt.lex_sign.astype(float)
    Mrt = E.mixed_mass_matrix(qd, rt_interp, E.HDIV, idx_r, ori_r, rt.ndofs, rt_interp, E.HDIV, idx_r, ori_r, rt.ndofs, I3)
    F = E.mixed_mass_matrix(qd, interp, E.HCURL, idx_n, ori_n, nd.ndofs, rt_interp, E.HDIV, idx_r, ori_r, rt.ndofs, I3)   # ND -> RT
    lu_rt, lu_nd = spla.splu(Mrt.tocsc()), spla.splu(c["M"].tocsc())
    se, ise = [np.sqrt(eps) * np.eye(3)] * ne, [np.eye(3) / np.sqrt(eps)] * ne
    acc = np.zeros(ne)
    for v, lam in zip(c["modes"], c["lam"]):
        omega = np.sqrt(lam / eps)                       # K v = lam M v with lam = omega^2 eps (mesh units, mu = 1)
        B = curl_dofs(nd, rt, v) / omega                 # |B| of B = -1 / (i omega) curl E
        Et = 0.5 * eps * (v @ (c["M"] @ v)) + 0.5 * (B @ (Mrt @ B))
        D = lu_rt.solve(eps * (F @ v))                   # eps E projected onto RT (GradFluxErrorEstimator)
        H = lu_nd.solve(F.T @ B)                         # mu^-1 B projected onto ND (CurlFluxErrorEstimator)
        eg = E.element_errors(qd, interp, E.HCURL, idx_n, ori_n, v, se, rt_interp, E.HDIV, idx_r, ori_r, D, ise)
        ec = E.element_errors(qd, rt_interp, E.HDIV, idx_r, ori_r, B, I3, interp, E.HCURL, idx_n, ori_n, H, I3)
        acc += 0.5 / Et * (eg + ec)
    return np.sqrt(acc / len(c["modes"]))


def test_error_indicator_statistics_match_the_reference():
    e = oracle_indicators(cavity_modes(int(FIX["order"])))
    print("Norm, Min, Max, Mean:", np.linalg.norm(e), e.min(), e.max(), e.mean())
    assert abs(np.linalg.norm(e) / REF_NORM - 1) < 1e-6
    assert abs(e.mean() / REF_MEAN - 1) < 2e-3 and abs(e.min() / REF_MIN - 1) < 2e-2 and abs(e.max() / REF_MAX - 1) < 5e-2
