"""TEST INFRASTRUCTURE: NumPy restatement of the flux error estimator of the reference for H(curl) problems, on explicitly
assembled matrices (no product code is used here; only tests import this module).

  FluxProjector                         /root/reference/palace/linalg/errorestimator.cpp:112-175
  ComputeErrorEstimates                 errorestimator.cpp:180-270
  CurlFluxErrorEstimator                errorestimator.cpp:400-513
  mixed H(div) -> H(curl) mass          /root/reference/palace/fem/qfunctions/33/hcurlhdiv_33_qf.h:33-55 (f_apply_hdivhcurl_33)
  element error integrand               /root/reference/palace/fem/qfunctions/33/hcurlhdiv_error_33_qf.h:46-76
  summation over the element            /root/reference/palace/fem/libceed/integrator.cpp:560-574 (all-ones basis)

Parity: (1) the pointwise arithmetic (mixed mass D, both element error integrands) is pinned to the reference's own QFunction
headers compiled in place (oracle/_ref, oracle/ref_qf.cpp) through tests/golden/qf_mixed_golden.npz
(tests/test_oracle_golden.py::test_estimator_oracle_matches_reference_golden, 1e-14); (2) END TO END: the statistics of the element
indicators the reference's regression suite stores for examples/cylinder/cavity_pec.json
(test/data/regression/ref/cylinder/cavity_pec/error-indicators.csv: 15 modes, grad-flux + curl-flux estimators, energy normalisation)
are reproduced by this module on the oracle-side discretisation -- the global norm to 1e-7, the reference's projection tolerance
being 1e-6 (tests/test_cylinder_indicator_golden.py); on tetrahedra the grad-flux estimator of examples/spheres (14,362 curved cubic
tets, RT_2) to 2.9e-8 in the global norm (tests/test_spheres_golden.py) and the complex Floquet chain of examples/cylinder/floquet.json
(minimum 1.7e-6, maximum 1.8e-5, norm 3e-4, per-mode energy defects 2e-3 relative: tests/test_cylinder_floquet_indicator_golden.py)."""
import numpy as np
import scipy.sparse as sp

HCURL, HDIV = 1, 2


def _mat33(col_major9):
    """3x3 matrix from the reference's column-major storage "0 3 6 / 1 4 7 / 2 5 8" (utils_33_qf.h)."""
    return np.asarray(col_major9, dtype=float).reshape(3, 3).T


def piola(map_type, adjJt):
    """Physical value = piola @ reference value: H(curl): adj(J)^T / det J = J^-T; H(div): J / det J."""
    A = _mat33(adjJt)
    if map_type == HCURL:
        return A
    return np.linalg.inv(A).T * np.linalg.det(A)  # J = A^-T, det J = 1 / det A


def mixed_mass_matrix(qdata, interp_trial, map_trial, idx_trial, sign_trial, n_trial, interp_test, map_test, idx_test, sign_test, n_test,
                      coef_elem):
    """Sparse [n_test x n_trial]: sum_e E_test^T B_test^T D B_trial E_trial, D = w detJ P_test^T C P_trial."""
    ne, _, Q = qdata.shape
    Pt, Ps = interp_trial.shape[2], interp_test.shape[2]
    rows, cols, vals = [], [], []
    for e in range(ne):
        Ae = np.zeros((Ps, Pt))
        for q in range(Q):
            w = qdata[e, 1, q]
            M1 = piola(map_trial, qdata[e, 2:, q])
            M2 = piola(map_test, qdata[e, 2:, q])
            D = w * M2.T @ coef_elem[e] @ M1
            Ae += interp_test[:, q, :].T @ D @ interp_trial[:, q, :]
        Ae = sign_test[e][:, None] * Ae * sign_trial[e][None, :]
        rows.append(np.repeat(idx_test[e], Pt))
        cols.append(np.tile(idx_trial[e], Ps))
        vals.append(Ae.ravel())
    return sp.coo_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(n_test, n_trial)).tocsr()


def element_errors(qdata, interp1, map1, idx1, sign1, x1, coef1_elem, interp2, map2, idx2, sign2, x2, coef2_elem):
    """eta_K^2 = sum_q w detJ | C2 P2 u2 - C1 P1 u1 |^2 per element."""
    ne, _, Q = qdata.shape
    out = np.zeros(ne)
    for e in range(ne):
        u1 = np.einsum("cqp,p->cq", interp1, sign1[e] * x1[idx1[e]])
        u2 = np.einsum("cqp,p->cq", interp2, sign2[e] * x2[idx2[e]])
        for q in range(Q):
            v1 = coef1_elem[e] @ (piola(map1, qdata[e, 2:, q]) @ u1[:, q])
            v2 = coef2_elem[e] @ (piola(map2, qdata[e, 2:, q]) @ u2[:, q])
            d = v2 - v1
            out[e] += qdata[e, 1, q] * (d @ d)
    return out


def spd_power(C, power):
    """MatrixSqrt / MatrixPow of a symmetric positive definite 3x3 (linalg/densematrix.cpp)."""
    w, V = np.linalg.eigh(np.asarray(C, dtype=float))
    return (V * w ** power) @ V.T


def rt_hex_tables(p, q1d):
    """interp[3][Q][P] of the RT_{p-1} hexahedron basis in the lexicographic order of the host layer (component c closed along
    axis c, open along the others; x fastest), built from the ORACLE's 1-D tables (independent of palace_b200/host)."""
    from oracle import pyoracle as O

    Bo, Bc, _, _ = O.nd_hex_1d(p, q1d)
    Q = q1d ** 3
    cols = []
    for c in range(3):
        n = [p, p, p]
        n[c] = p + 1
        for k in range(n[2]):
            for j in range(n[1]):
                for i in range(n[0]):
                    tx = Bc[:, i] if c == 0 else Bo[:, i]
                    ty = Bc[:, j] if c == 1 else Bo[:, j]
                    tz = Bc[:, k] if c == 2 else Bo[:, k]
                    v = np.zeros((3, Q))
                    v[c] = np.einsum("c,b,a->cba", tz, ty, tx).ravel()
                    cols.append(v)
    return np.stack(cols, axis=2)
