"""CPU: the host layer's order-1 ND prism and mixed hexahedron / prism meshes (palace_b200/host/prism.py) -- the second element
geometry of a BilinearForm (/root/reference/palace/fem/bilinearform.cpp:56-101: one sub-operator per geometry type). MFEM's
ND_WedgeElement is not in /root/reference, so the element is held to what it must do: its curl table is the curl of its value table,
edge circulations are the dofs, and the PEC cavity eigenvalues of a box meshed half with hexahedra, half with prisms converge to
the closed form at second order with the null space exactly the gradients."""
import numpy as np
import scipy.linalg as sla

from oracle import pyoracle as O
from oracle import solvers as S
from palace_b200.host import hexspace as hs
from palace_b200.host import prism as pr
from tests import common


def test_prism_tables_are_consistent():
    pts, w = pr.prism_quadrature()
    assert abs(w.sum() - 0.5) < 1e-15
    interp, curl = pr.prism_tables(pts)
    eps = 1e-6
    for d in range(3):  # curl by central differences of the value table
        dp, dm = pts.copy(), pts.copy()
        dp[:, d] += eps
        dm[:, d] -= eps
        Ip, _ = pr.prism_tables(dp)
        Im, _ = pr.prism_tables(dm)
        if d == 0:
            grad = np.zeros((3, 3) + interp.shape[1:])
        grad[:, d] = (Ip - Im) / (2 * eps)  # grad[c, d] = d phi_c / d x_d
    num = np.stack([grad[2, 1] - grad[1, 2], grad[0, 2] - grad[2, 0], grad[1, 0] - grad[0, 1]])
    assert np.abs(num - curl).max() < 1e-9
    # circulation of function j along edge i is delta_ij (midpoint rule is exact: the tangential component is constant on an edge)
    V = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1], [1, 0, 1], [0, 1, 1]], dtype=float)
    mids = np.array([(V[a] + V[b]) / 2 for a, b in pr.PRISM_EDGES])
    Im, _ = pr.prism_tables(mids)
    C = np.array([[(V[b] - V[a]) @ Im[:, i, j] for j in range(9)] for i, (a, b) in enumerate(pr.PRISM_EDGES)])
    assert np.abs(C - np.eye(9)).max() < 1e-14


def mixed_matrices(n, nxh):
    mesh = pr.mixed_box_mesh(n, nxh)
    sp_ = pr.build_mixed_nd_space(mesh)
    out = {}
    nodes = hs.gauss_lobatto(2)
    xe = mesh.hexes.node_coords(1, nodes)
    qd_h = O.geom_hex_qdata(xe, mesh.hexes.attr, 1, 2)
    ih, ch, _ = O.nd_hex_tables(1, 2)
    idx_h, ori_h = sp_.hex_space.native_restriction()
    pts, w = pr.prism_quadrature()
    ip, cp_ = pr.prism_tables(pts)
    qd_p = pr.prism_qdata(mesh, w)
    for kind in (O.CURLCURL, O.ND_MASS):
        blob = common.coefficient(kind, 1, "const")
        Ah = S.assemble_sparse(O.element_matrices(kind, ih, ch, ori_h, qd_h, blob, 12), idx_h.astype(np.int64), sp_.ndofs)
        Ap = S.assemble_sparse(O.element_matrices(kind, ip, cp_, sp_.prism_orient, qd_p, blob, 9), sp_.prism_idx.astype(np.int64), sp_.ndofs)
        out[kind] = (Ah + Ap).tocsr()
    return mesh, sp_, out


def cavity_eigenvalues(n, nxh, k=3):
    _, sp_, A = mixed_matrices(n, nxh)
    free = np.setdiff1d(np.arange(sp_.ndofs), sp_.ess_dofs)
    K, M = A[O.CURLCURL][free][:, free].toarray(), A[O.ND_MASS][free][:, free].toarray()
    w = sla.eigh(K, M, eigvals_only=True)
    return w[w > 1.0][:k], int((w < 1e-8).sum())


def test_mixed_mesh_cavity_eigenvalues_converge_at_second_order():
    exact = np.pi ** 2 * np.array([1.25, 1.25, 2.0])  # box 2 x 1 x 1
    w1, null1 = cavity_eigenvalues((4, 2, 2), 2)
    w2, null2 = cavity_eigenvalues((8, 4, 4), 4)
    e1, e2 = np.abs(w1 - exact) / exact, np.abs(w2 - exact) / exact
    assert e2.max() < 0.07 and np.all(e1 / e2 > 3.0) and np.all(e1 / e2 < 5.5), (e1, e2)
    # null space = gradients of the interior vertices: (nx - 1)(ny - 1)(nz - 1)
    assert null1 == 3 * 1 * 1 and null2 == 7 * 3 * 3
