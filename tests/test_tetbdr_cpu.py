"""CPU: boundary mass on tetrahedral meshes through the parent element's tables at the face points
(palace_b200/host/tetbdr.py; the reference's VectorFEMassIntegrator on boundary triangles, spaceoperator.cpp:300-303): exact surface
integrals of polynomial fields, symmetry, and no coupling to dofs away from the boundary."""
import numpy as np
import pytest

from oracle import pyoracle as O
from palace_b200.host import coeff as cf
from palace_b200.host import tetbdr as tb
from palace_b200.host import tetspace as ts


def bdr_apply(groups, blob, x, n):
    y = np.zeros(n)
    for g in groups:
        O.apply_add_co(O.ND_MASS, g.interp, None, g.idx, g.curl_orient, g.qdata, blob, x, y)
    return y


def test_triangle_rule():
    for deg in (1, 4, 7):
        pts, w = tb.tri_quadrature(deg)
        assert abs(w.sum() - 0.5) < 1e-15 and (pts >= 0).all() and (pts.sum(axis=1) <= 1).all()
        from math import factorial

        for a in range(deg + 1):
            b = deg - a
            assert abs((w * pts[:, 0] ** a * pts[:, 1] ** b).sum() - factorial(a) * factorial(b) / factorial(a + b + 2)) < 1e-15


@pytest.mark.parametrize("p", [1, 2, 3])
def test_boundary_mass_energy_of_polynomial_fields(p):
    size = (1.0, 0.8, 0.9)
    mesh = ts.box_tet_mesh((2, 2, 1), size, jitter=0.2, scramble_seed=3)
    nd = ts.build_nd_tet_space(mesh, p)
    groups = tb.boundary_groups(mesh, nd)
    assert abs(sum(g.areas.sum() for g in groups) - 2 * (size[0] * size[1] + size[0] * size[2] + size[1] * size[2])) < 1e-12
    am, mats = tb.tangential_materials(groups, c=1.7)
    blob = cf.coeff_ctx(am, mats)
    a = np.array([0.7, -1.1, 0.4])
    x = ts.interpolate(mesh, nd, lambda X: a)
    y = bdr_apply(groups, blob, x, nd.ndofs)
    want = 1.7 * sum(2 * area * (a @ a - a[d] ** 2) for d, area in ((0, size[1] * size[2]), (1, size[0] * size[2]), (2, size[0] * size[1])))
    assert abs(x @ y - want) < 1e-11 * want
    if p >= 2:  # a linear field E = (y, z, x): |E_t|^2 integrated over the six faces of the box [0, a] x [0, b] x [0, c]
        x2 = ts.interpolate(mesh, nd, lambda X: np.array([X[1], X[2], X[0]]))
        y2 = bdr_apply(groups, blob, x2, nd.ndofs)
        a_, b_, c_ = size
        fx = lambda x0: b_ * c_ ** 3 / 3 + x0 ** 2 * b_ * c_          # x = x0: (E_y, E_z) = (z, x0)
        fy = lambda y0: y0 ** 2 * a_ * c_ + a_ ** 3 * c_ / 3          # y = y0: (E_x, E_z) = (y0, x)
        fz = lambda z0: a_ * b_ ** 3 / 3 + z0 ** 2 * a_ * b_          # z = z0: (E_x, E_y) = (y, z0)
        want2 = 1.7 * (fx(0.0) + fx(a_) + fy(0.0) + fy(b_) + fz(0.0) + fz(c_))
        assert abs(x2 @ y2 - want2) < 1e-10 * want2
    # symmetric, positive semi-definite, and blind to dofs off the boundary
    z = np.random.default_rng(1).random(nd.ndofs)
    yz = bdr_apply(groups, blob, z, nd.ndofs)
    interior = np.setdiff1d(np.arange(nd.ndofs), nd.ess_dofs)
    assert np.abs(yz[interior]).max() < 1e-13 * np.abs(yz).max() and z @ yz > 0
    assert abs(x @ yz - z @ y) < 1e-12 * abs(z @ yz)
