"""End-to-end pin of the FLOQUET-periodic terms against the reference's own output (CPU, oracle side): the reference's
regression suite stores the eigenfrequencies of examples/cylinder/floquet.json -- the tetrahedral waveguide of
test_cylinder_tet_golden.py (cylinder_tet.msh, Nedelec order 4, periodic in z, PEC wall, lossy Teflon) with the Floquet wave
vector k = (0, 0, 0.4) -- in test/data/regression/ref/cylinder/floquet/eig.csv. With [k x] the cross-product matrix the reference
assembles (/root/reference/palace/models/spaceoperator.cpp:370-395,1265-1283, materialoperator.cpp:365-373)

    K = curl-curl(mu^-1) + mass([k x]^T mu^-1 [k x])  +  i ( MixedVectorWeakCurl(mu^-1 [k x]) + MixedVectorCurl(mu^-1 [k x], transpose) )

i.e. the two mixed curl integrators of fem/integ/mixedveccurl.cpp carry the imaginary part. The oracle's kinds ND_WEAKCURL /
ND_MIXEDCURL (pointwise arithmetic pinned to the reference's hcurlhdiv_33_qf.h in test_oracle_golden.py) assembled on the
oracle-side tet space reproduce the stored frequencies; a sign, transpose or map-type error in either term moves them at the
1e-2 level (the test also solves the k = 0 problem to show the distance)."""
import os

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from oracle import pyoracle as O
from palace_b200.host import coeff as cf
from tests import common
from tests.test_cylinder_tet_golden import FIX, frequencies_ghz, sigma_target, space_and_tables


def cross_matrix(k):
    return np.array([[0.0, -k[2], k[1]], [k[2], 0.0, -k[0]], [-k[1], k[0], 0.0]])


def assembled(nd, kind, interp, curl, qd, blob):
    Ae = O.element_matrices(kind, interp, curl, None, qd, blob, nd.P)
    rows, cols, vals = [], [], []
    for e in range(Ae.shape[0]):
        T = nd.dense_T(e)
        r, c = np.meshgrid(nd.idx[e], nd.idx[e], indexing="ij")
        rows.append(r.ravel())
        cols.append(c.ravel())
        vals.append((T.T @ Ae[e] @ T).ravel())
    return sp.csr_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(nd.ndofs, nd.ndofs))


def floquet_matrices(nd, interp, curl, qd, k):
    """(K_r, K_i, M) of the reference's Floquet formulation with unit mu^-1 and unit mass coefficient."""
    kx = cross_matrix(k)
    am = np.array([0])
    one = cf.coeff_ctx(a=1.0)
    Kr = assembled(nd, O.CURLCURL, interp, curl, qd, one) + assembled(nd, O.ND_MASS, interp, curl, qd, cf.coeff_ctx(am, (kx.T @ kx)[None]))
    Ki = (assembled(nd, O.ND_WEAKCURL, interp, curl, qd, cf.coeff_ctx(am, kx[None], a=-1.0))              # -(F u, curl v)
          + assembled(nd, O.ND_MIXEDCURL, interp, curl, qd, cf.coeff_ctx(am, kx[None], transpose=True)))   # (F^T curl u, v)
    M = assembled(nd, O.ND_MASS, interp, curl, qd, one)
    return Kr.tocsr(), Ki.tocsr(), M.tocsr()


def test_floquet_eigenfrequencies_match_the_reference():
    p = int(FIX["order"])
    mesh, nd, interp, curl, qd = space_and_tables(p)
    k = FIX["floquet_wave_vector"]
    Kr, Ki, M = floquet_matrices(nd, interp, curl, qd, k)
    assert abs(Ki + Ki.T).max() < 1e-12 * abs(Ki).max()          # the pair is skew-symmetric: K_r + i K_i is Hermitian
    free = np.setdiff1d(np.arange(nd.ndofs), nd.ess_dofs)
    # the Hermitian pencil (K_r + i K_i) z = lam M z as the real symmetric one of twice the size, [[K_r, -K_i], [K_i, K_r]] (x; y) =
    # lam diag(M, M) (x; y) with z = x + i y: every eigenvalue appears twice (SciPy's ARPACK wrapper mishandles complex shift-invert
    # with a mass matrix here)
    Krf, Kif, Mf = Kr[free][:, free], Ki[free][:, free], M[free][:, free]
    A2 = sp.bmat([[Krf, -Kif], [Kif, Krf]], format="csc")
    M2 = sp.block_diag([Mf, Mf], format="csc")
    lam2 = common.eigsh_above(A2, M2, 30, sigma_target(), extra=8)
    assert np.abs(lam2[0::2] - lam2[1::2]).max() < 1e-9 * lam2.max()
    lam = 0.5 * (lam2[0::2] + lam2[1::2])
    f = frequencies_ghz(lam)
    rel = np.abs(f.real - FIX["floquet_f_re_ghz"]) / FIX["floquet_f_re_ghz"]
    print("Re f (GHz), oracle tet space with the Floquet terms:", f.real)
    print("rel. error vs the reference's floquet/eig.csv:", rel)
    assert rel.max() < 5e-8 and np.sum(rel < 1e-8) >= 10
    assert (np.abs(f.imag - FIX["floquet_f_im_ghz"]) / FIX["floquet_f_im_ghz"]).max() < 1e-6
    # the periodic terms matter: the k = 0 spectrum (test_cylinder_tet_golden.py) sits far from these numbers
    assert (np.abs(FIX["ref_f_re_ghz"] - FIX["floquet_f_re_ghz"]) / FIX["floquet_f_re_ghz"]).max() > 1e-2
