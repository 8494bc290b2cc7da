"""Caller-side stand-in: reader for the reference's example meshes in Gmsh 2.2 binary format with second-order
hexahedra (HEX27 volume elements + QUAD9 boundary elements, e.g. /root/reference/examples/cylinder/mesh/cylinder_hex.msh),
converted to the lexicographic order-2 node layout the geometry entry points take (b2p_geom_create_hex with mesh_order 2:
27 nodes per element, x fastest, nodes at 0, 1/2, 1 = the three Gauss-Lobatto points). MFEM does this conversion inside
Palace (fem/mesh.cpp); here it only serves the end-to-end checks against the reference's stored regression outputs."""
from __future__ import annotations

import dataclasses
import struct

import numpy as np

# Gmsh corner numbering of a hexahedron -> (x, y, z) in {0, 1}^3
_GMSH_CORNERS = [(0, 0, 0), (1, 0, 0), (1, 1, 0), (0, 1, 0), (0, 0, 1), (1, 0, 1), (1, 1, 1), (0, 1, 1)]
_GMSH_EDGES = [(0, 1), (0, 3), (0, 4), (1, 2), (1, 5), (2, 3), (2, 6), (3, 7), (4, 5), (4, 7), (5, 6), (6, 7)]
_GMSH_FACES = [(0, 3, 2, 1), (0, 1, 5, 4), (0, 4, 7, 3), (1, 2, 6, 5), (2, 3, 7, 6), (4, 5, 6, 7)]
_NODES_PER_TYPE = {1: 2, 2: 3, 3: 4, 4: 4, 5: 8, 8: 3, 9: 6, 10: 9, 11: 10, 12: 27, 15: 1}


def hex27_reference_positions():
    """Reference coordinates (multiples of 1/2) of Gmsh's 27 hexahedron nodes, in Gmsh order."""
    c = np.array(_GMSH_CORNERS, dtype=np.float64)
    pos = [c[i] for i in range(8)]
    pos += [0.5 * (c[a] + c[b]) for a, b in _GMSH_EDGES]
    pos += [0.25 * (c[a] + c[b] + c[d] + c[e]) for a, b, d, e in _GMSH_FACES]
    pos.append(np.array([0.5, 0.5, 0.5]))
    return np.array(pos)


@dataclasses.dataclass
class Hex27Mesh:
    verts: np.ndarray     # [nv][3] corner vertices actually used (renumbered)
    elems: np.ndarray     # [ne][8] corner vertex ids, corner = x + 2 y + 4 z
    attr: np.ndarray      # [ne] volume attribute (Gmsh physical tag)
    xe2: np.ndarray       # [ne][3][27] order-2 nodes, lexicographic (x fastest), component-major
    bdr_attr: np.ndarray  # [nb] attribute of the QUAD9 boundary elements
    bdr_verts: np.ndarray  # [nb][4] corner vertex ids of the boundary elements

    @property
    def ne(self):
        return int(self.elems.shape[0])


def read_gmsh22_binary(path: str):
    """(nodes {id: xyz}, elements {type: (tags[n][ntags], conn[n][nn])}) of a Gmsh 2.2 binary file."""
    f = open(path, "rb").read()
    assert f[f.index(b"$MeshFormat\n") + 12:].split(b"\n", 1)[0].split()[:3] == [b"2.2", b"1", b"8"], "Gmsh 2.2 binary, 8-byte reals"
    i = f.index(b"$Nodes\n") + 7
    j = f.index(b"\n", i)
    n = int(f[i:j])
    rec = np.dtype([("id", "<i4"), ("x", "<f8", 3)])
    nodes = np.frombuffer(f, dtype=rec, count=n, offset=j + 1)
    off = j + 1 + n * rec.itemsize
    k = f.index(b"$Elements\n", off) + 10
    j = f.index(b"\n", k)
    ne = int(f[k:j])
    off = j + 1
    out, cnt = {}, 0
    while cnt < ne:
        etype, nfollow, ntags = struct.unpack_from("<3i", f, off)
        off += 12
        nn = _NODES_PER_TYPE[etype]
        w = 1 + ntags + nn
        blk = np.frombuffer(f, dtype="<i4", count=w * nfollow, offset=off).reshape(nfollow, w)
        off += 4 * w * nfollow
        cnt += nfollow
        tags, conn = out.setdefault(etype, ([], []))
        tags.append(blk[:, 1:1 + ntags])
        conn.append(blk[:, 1 + ntags:])
    return {int(r["id"]): r["x"].copy() for r in nodes}, {t: (np.concatenate(a), np.concatenate(b)) for t, (a, b) in out.items()}


def load_hex27(path: str) -> Hex27Mesh:
    nodes, elems = read_gmsh22_binary(path)
    tags, conn = elems[12]
    ref = hex27_reference_positions()
    lex = (np.rint(2 * ref[:, 0]) + 3 * np.rint(2 * ref[:, 1]) + 9 * np.rint(2 * ref[:, 2])).astype(int)  # Gmsh -> lexicographic
    ne = conn.shape[0]
    xe2 = np.zeros((ne, 3, 27))
    for e in range(ne):
        X = np.array([nodes[int(v)] for v in conn[e]])                      # Gmsh order
        # sanity of the ordering table: every mid node sits near the mean of the corners it belongs to
        c8 = X[:8]
        w = np.array([[(1 - r[0] if cx == 0 else r[0]) * (1 - r[1] if cy == 0 else r[1]) * (1 - r[2] if cz == 0 else r[2])
                       for (cx, cy, cz) in _GMSH_CORNERS] for r in ref])
        size = np.linalg.norm(c8.max(axis=0) - c8.min(axis=0))
        assert np.linalg.norm(w @ c8 - X, axis=1).max() < 0.2 * size, "HEX27 node ordering does not match the Gmsh convention"
        xe2[e][:, lex] = X.T
    corner_ids = np.unique(conn[:, :8])
    renum = {int(v): i for i, v in enumerate(corner_ids)}
    perm = [0, 1, 3, 2, 4, 5, 7, 6]  # ours (x + 2y + 4z) <- Gmsh corner number
    el = np.array([[renum[int(conn[e, perm[c]])] for c in range(8)] for e in range(ne)], dtype=np.int64)
    verts = np.array([nodes[int(v)] for v in corner_ids])
    bt, bc = elems.get(10, (np.zeros((0, 2), dtype=int), np.zeros((0, 9), dtype=int)))
    bverts = np.array([[renum[int(v)] for v in row[:4]] for row in bc], dtype=np.int64).reshape(-1, 4)
    return Hex27Mesh(verts, el, tags[:, 0].astype(np.int32), xe2, bt[:, 0].astype(np.int32) if len(bt) else np.zeros(0, np.int32), bverts)


# ------------------------------------------------------------------------------------------------
# High-order tetrahedra (TET10 / TET20 volume elements with TRI6 / TRI10 boundary elements)
# ------------------------------------------------------------------------------------------------
_NODES_PER_TYPE.update({21: 10, 29: 20, 26: 4})
_TET_TYPES = {4: 1, 11: 2, 29: 3}
_TRI_TYPES = {2: 1, 9: 2, 21: 3}


@dataclasses.dataclass
class HighOrderTetMesh:
    order: int
    verts: np.ndarray      # [nv][3] corner vertices (renumbered)
    elems: np.ndarray      # [ne][4] corner vertex ids, positively oriented as stored
    attr: np.ndarray       # [ne]
    xe: np.ndarray         # [ne][3][Nn] element nodes on the equispaced lattice (a fastest, a + b + c <= order), component-major
    bdr_attr: np.ndarray   # [nb]
    bdr_verts: np.ndarray  # [nb][3] corner vertex ids of the boundary triangles

    @property
    def ne(self):
        return int(self.elems.shape[0])


def load_tets(path: str) -> HighOrderTetMesh:
    """Order-k tetrahedral mesh; the high-order nodes of every element are matched to the lattice positions geometrically
    (nearest node to the affine image of the lattice point), so no table of Gmsh's node ordering is involved."""
    nodes, elems = read_gmsh22_binary(path)
    (etype,) = [t for t in elems if t in _TET_TYPES]
    order = _TET_TYPES[etype]
    tags, conn = elems[etype]
    lat = np.array([(a, b, c) for c in range(order + 1) for b in range(order + 1 - c) for a in range(order + 1 - b - c)], dtype=np.float64) / order
    ids = np.array(sorted(nodes))
    pos = np.zeros((ids.max() + 1, 3))
    pos[ids] = np.array([nodes[int(i)] for i in ids])
    X = pos[conn]                                                  # [ne][Nn][3] in Gmsh order
    c0 = X[:, 0]
    E = np.stack([X[:, 1] - c0, X[:, 2] - c0, X[:, 3] - c0], axis=1)   # [ne][3 edges][3]
    img = c0[:, None, :] + np.einsum("nd,edc->enc", lat, E)        # affine images of the lattice points
    d = np.linalg.norm(img[:, :, None, :] - X[:, None, :, :], axis=3)  # [ne][lattice][gmsh node]
    pick = d.argmin(axis=2)
    h = np.linalg.norm(E, axis=2).min(axis=1)
    assert (np.sort(pick, axis=1) == np.arange(lat.shape[0])[None]).all(), "lattice matching is not a permutation"
    # curved elements: a node may sit off its affine image, but never by as much as the lattice spacing
    assert (np.take_along_axis(d, pick[:, :, None], axis=2)[:, :, 0].max(axis=1) < 0.45 * np.linalg.norm(E, axis=2).max(axis=1) / order).all()
    xe = np.transpose(np.take_along_axis(X, pick[:, :, None], axis=1), (0, 2, 1))
    corner_ids = np.unique(conn[:, :4])
    renum = np.full(ids.max() + 1, -1, dtype=np.int64)
    renum[corner_ids] = np.arange(corner_ids.size)
    (btype,) = [t for t in elems if t in _TRI_TYPES] or [None]
    bt, bc = elems[btype] if btype is not None else (np.zeros((0, 2), dtype=int), np.zeros((0, 3), dtype=int))
    return HighOrderTetMesh(order, pos[corner_ids], renum[conn[:, :4]], tags[:, 0].astype(np.int32), np.ascontiguousarray(xe),
                            bt[:, 0].astype(np.int32), renum[bc[:, :3]])


def refine_hex27(mesh: Hex27Mesh) -> Hex27Mesh:
    """One uniform refinement (Model.Refinement.UniformLevels): every hexahedron becomes eight, the children's order-2 nodes
    are the parent's triquadratic map evaluated at the child node positions (what refining MFEM's nodal grid function does)."""
    t = np.array([0.0, 0.5, 1.0])

    def lag(x):  # quadratic Lagrange basis on {0, 1/2, 1} at points x: [len(x)][3]
        return np.stack([2 * (x - 0.5) * (x - 1), -4 * x * (x - 1), 2 * x * (x - 0.5)], axis=1)

    ne = mesh.ne
    X = mesh.xe2.reshape(ne, 3, 3, 3, 3)                        # [e][c][k][j][i]
    kids, attr = [], []
    for oz in (0, 1):
        for oy in (0, 1):
            for ox in (0, 1):
                Lx, Ly, Lz = lag(0.5 * (t + ox)), lag(0.5 * (t + oy)), lag(0.5 * (t + oz))
                kids.append(np.einsum("eckji,ai,bj,dk->ecdba", X, Lx, Ly, Lz).reshape(ne, 3, 27))
                attr.append(mesh.attr)
    xe2 = np.concatenate(kids, axis=0)
    # corner vertices: unique points among the children's 8 corners
    cidx = [0, 2, 6, 8, 18, 20, 24, 26]
    corners = np.transpose(xe2[:, :, cidx], (0, 2, 1)).reshape(-1, 3)
    scale = np.abs(corners).max()
    key = np.rint(corners / scale * 2 ** 40).astype(np.int64)
    _, first, inv = np.unique(key, axis=0, return_index=True, return_inverse=True)
    return Hex27Mesh(corners[first], inv.reshape(-1, 8), np.concatenate(attr), np.ascontiguousarray(xe2), mesh.bdr_attr[:0], np.zeros((0, 4), dtype=np.int64))
