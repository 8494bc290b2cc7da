"""Lowest-order Nedelec prisms (wedges) and mixed hexahedron / prism meshes: caller-side stand-in for the second geometry type
of a Palace mesh. ``BilinearForm::PartialAssemble`` creates one sub-operator per element geometry type
(/root/reference/palace/fem/bilinearform.cpp:56-101); every non-tensor vector element reaches libCEED as dense tables
(/root/reference/palace/fem/libceed/basis.cpp:40-85), which is what ``b2p_op_create_dense`` takes, so a prism block needs nothing
from the library but its tables, q-data and restriction. MFEM (``ND_WedgeElement``) is not in /root/reference; the element here is
the classical first-kind lowest-order one with edge-circulation dofs,

    horizontal edges (a, b) of the bottom / top triangle:  W_ab (1 - z),  W_ab z,   W_ab = l_a grad l_b - l_b grad l_a
    vertical edges a:                                       l_a e_z

(l_a the barycentric coordinates of the triangle), checked by its cavity eigenvalues on mixed meshes
(tests/test_prism_cpu.py). Order 1 only: higher orders bring face dofs on triangles, whose orientation handling the dense
path has (the int8 tridiagonal rows of restriction.cpp:301-368) but this host layer does not generate."""
from __future__ import annotations

import dataclasses

import numpy as np

from . import hexmesh as hm
from . import hexspace as hs

# local edges as (vertex a, vertex b), oriented a -> b; vertices 0, 1, 2 bottom triangle, 3, 4, 5 top (MFEM's wedge numbering)
PRISM_EDGES = [(0, 1), (1, 2), (0, 2), (3, 4), (4, 5), (3, 5), (0, 3), (1, 4), (2, 5)]
_TRI_PTS = np.array([[1 / 6, 1 / 6], [2 / 3, 1 / 6], [1 / 6, 2 / 3]])  # degree-2 rule on the reference triangle
_TRI_W = np.array([1 / 6, 1 / 6, 1 / 6])


def prism_quadrature():
    """6 points: the 3-point triangle rule x 2 Gauss-Legendre points in z (exact for the order-1 mass and curl-curl forms on
    affine prisms)."""
    z, wz = hs.gauss_legendre(2)
    pts = np.array([[x, y, zz] for zz in z for (x, y) in _TRI_PTS])
    w = np.array([wt * wzz for wzz in wz for wt in _TRI_W])
    return pts, w


def prism_tables(pts):
    """interp[3][Q][9], curl[3][Q][9] of the order-1 ND prism in reference coordinates."""
    Q = len(pts)
    interp = np.zeros((3, Q, 9))
    curl = np.zeros((3, Q, 9))
    gl = np.array([[-1.0, -1.0], [1.0, 0.0], [0.0, 1.0]])  # grad l_a
    for q, (x, y, z) in enumerate(pts):
        lam = np.array([1.0 - x - y, x, y])
        for j, (a, b) in enumerate(PRISM_EDGES[:6]):
            a, b = a % 3, b % 3
            W = lam[a] * gl[b] - lam[b] * gl[a]
            cz = 2.0 * (gl[a][0] * gl[b][1] - gl[a][1] * gl[b][0])
            g, dg = (1.0 - z, -1.0) if j < 3 else (z, 1.0)
            interp[0, q, j], interp[1, q, j] = W[0] * g, W[1] * g
            curl[0, q, j], curl[1, q, j], curl[2, q, j] = -W[1] * dg, W[0] * dg, g * cz
        for a in range(3):
            j = 6 + a
            interp[2, q, j] = lam[a]
            curl[0, q, j], curl[1, q, j] = gl[a][1], -gl[a][0]
    return interp, curl


@dataclasses.dataclass
class MixedMesh:
    verts: np.ndarray       # [NV, 3]
    hexes: hm.HexMesh       # the hexahedral block (global vertex ids)
    prisms: np.ndarray      # [NP, 6] global vertex ids, MFEM wedge order
    prism_attr: np.ndarray  # [NP] int32
    size: tuple


def mixed_box_mesh(n, nxh, h=None, n_attr=1) -> MixedMesh:
    """Box of n = (nx, ny, nz) cells: the first ``nxh`` layers along x stay hexahedra, every other cell is cut into two prisms
    along the diagonal of its x-y face (triangles extruded in z)."""
    nx, ny, nz = n
    h = 1.0 / ny if h is None else h
    full = hm.box_mesh((nx, ny, nz), size=(nx * h, ny * h, nz * h))
    ci = np.arange(full.ne) % nx  # x index of the cell (box_mesh orders cells x fastest)
    hsel = ci < nxh
    hexes = hm.HexMesh(verts=full.verts, elems=full.elems[hsel], attr=(1 + np.arange(int(hsel.sum())) % n_attr).astype(np.int32))
    pr = []
    for el in full.elems[~hsel]:
        v = lambda a, b, c: el[a + 2 * b + 4 * c]
        pr.append([v(0, 0, 0), v(1, 0, 0), v(1, 1, 0), v(0, 0, 1), v(1, 0, 1), v(1, 1, 1)])
        pr.append([v(0, 0, 0), v(1, 1, 0), v(0, 1, 0), v(0, 0, 1), v(1, 1, 1), v(0, 1, 1)])
    prisms = np.array(pr, dtype=np.int64).reshape(-1, 6)
    return MixedMesh(full.verts, hexes, prisms, (1 + np.arange(len(prisms)) % n_attr).astype(np.int32), (nx * h, ny * h, nz * h))


def prism_qdata(mesh: MixedMesh, w):
    """q-data [NP][11][Q] of the affine prisms in the reference layout {attr, w detJ, J^-T column-major} (geom_33_qf.h:9-34)."""
    Q = len(w)
    qd = np.empty((len(mesh.prisms), 11, Q))
    for e, pv in enumerate(mesh.prisms):
        X = mesh.verts[pv]
        J = np.stack([X[1] - X[0], X[2] - X[0], X[3] - X[0]], axis=1)
        det = np.linalg.det(J)
        assert det > 0
        qd[e, 0] = mesh.prism_attr[e]
        qd[e, 1] = w * det
        qd[e, 2:] = np.linalg.inv(J).T.ravel(order="F")[:, None]
    return qd


@dataclasses.dataclass
class MixedNDSpace:
    """Order-1 ND space on a mixed mesh: one dof per edge, oriented from the lower to the higher global vertex."""
    ndofs: int
    hex_space: hs.HexSpace     # the hexahedral block with GLOBAL dof numbers
    prism_idx: np.ndarray      # [NP, 9] int32
    prism_orient: np.ndarray   # [NP, 9] int8
    ess_dofs: np.ndarray       # dofs on the boundary of the box


def build_mixed_nd_space(mesh: MixedMesh) -> MixedNDSpace:
    topo = hs.build_topology(mesh.hexes)
    hsp = hs.build_nd_space(mesh.hexes, topo, 1)  # p = 1: dof = edge id of the block's own topology
    edges = {}

    def eid(a, b):
        key = (min(a, b), max(a, b))
        if key not in edges:
            edges[key] = len(edges)
        return edges[key]

    hex_global = np.array([eid(int(a), int(b)) for a, b in topo.edge_verts], dtype=np.int64)
    pidx = np.empty((len(mesh.prisms), 9), dtype=np.int32)
    pori = np.empty((len(mesh.prisms), 9), dtype=np.int8)
    for e, pv in enumerate(mesh.prisms):
        for j, (a, b) in enumerate(PRISM_EDGES):
            ga, gb = int(pv[a]), int(pv[b])
            pidx[e, j] = eid(ga, gb)
            pori[e, j] = 1 if ga < gb else -1
    ndofs = len(edges)
    hsp = dataclasses.replace(hsp, ndofs=ndofs, lex_gid=hex_global[hsp.lex_gid], mult=np.zeros(ndofs, dtype=np.int64))
    # essential dofs: edges with both vertices on one boundary plane of the box
    ess = []
    L = np.array(mesh.size)
    for (a, b), g in edges.items():
        xa, xb = mesh.verts[a], mesh.verts[b]
        if any((abs(xa[d]) < 1e-12 and abs(xb[d]) < 1e-12) or (abs(xa[d] - L[d]) < 1e-12 and abs(xb[d] - L[d]) < 1e-12) for d in range(3)):
            ess.append(g)
    return MixedNDSpace(ndofs, hsp, pidx, pori, np.array(sorted(ess), dtype=np.int64))
