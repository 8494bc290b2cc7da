"""CPU checks of the complex Krylov restatement in oracle/solvers.py (the checker used by tests/test_complex_gpu.py):
every Gram-Schmidt / preconditioner-side variant must reach the direct solution of a non-Hermitian complex system,
and the complex Givens rotation must annihilate the second component (iterative.cpp:112-226)."""
import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from oracle import solvers as S


def test_complex_plane_rotation_annihilates():
    rng = np.random.default_rng(0)
    for _ in range(20):
        dx, dy = complex(*rng.standard_normal(2)), complex(*rng.standard_normal(2))
        cs, sn = S._cplane_rotation(dx, dy)
        a, b = S._capply_rot(dx, dy, cs, sn)
        assert abs(b) < 1e-14 and abs(abs(a) - np.hypot(abs(dx), abs(dy))) < 1e-14
        assert abs(cs * cs + abs(sn) ** 2 - 1.0) < 1e-14
    assert S._cplane_rotation(1.0 + 1j, 0.0) == (1.0, 0.0)
    cs, sn = S._cplane_rotation(0.0, 2.0j)
    assert cs == 0.0 and abs(sn * 2.0j - 2.0) < 1e-15


@pytest.mark.parametrize("orthog", [0, 1, 2])
@pytest.mark.parametrize("flexible,right", [(False, True), (False, False), (True, True)])
def test_complex_gmres_variants(orthog, flexible, right):
    n = 150
    rng = np.random.default_rng(1)
    A = (sp.random(n, n, 0.05, random_state=1) + 1j * sp.random(n, n, 0.05, random_state=2) + 6.0 * sp.eye(n)).tocsr()
    b = rng.standard_normal(n) + 1j * rng.standard_normal(n)
    d = A.diagonal()
    x, it, conv = S.cgmres(A, b, lambda r: r / d.real, rel_tol=1e-11, max_it=150, max_dim=25, orthog=orthog, flexible=flexible,
                           right=right)
    assert conv and it < 60
    assert np.linalg.norm(x - spla.spsolve(A.tocsc(), b)) < 1e-9 * np.linalg.norm(x)


def test_complex_gram_schmidt_conventions():
    rng = np.random.default_rng(2)
    n, m = 300, 7
    Q, _ = np.linalg.qr(rng.standard_normal((n, m)) + 1j * rng.standard_normal((n, m)))
    V = [np.ascontiguousarray(Q[:, j]) for j in range(m)]
    w = rng.standard_normal(n) + 1j * rng.standard_normal(n)
    for kind in (0, 1, 2):
        H, w2 = S.corthogonalize(kind, V, w)
        assert np.abs(H - Q.conj().T @ w).max() < 1e-12
        assert np.abs(Q.conj().T @ w2).max() < 1e-12
