"""Non-conforming (one level of hanging faces) hexahedral meshes and the conforming prolongation of their ND spaces: the
caller-side stand-in for what MFEM gives Palace on an AMR mesh, ``ParFiniteElementSpace::GetProlongationMatrix()`` -- the matrix
``ParOperator`` multiplies by before and after the local operator (/root/reference/palace/linalg/rap.cpp:195-234) and whose
entry-wise absolute value assembles the diagonal (rap.cpp:162-178).

The mesh is a box: ``nc`` coarse cells of size ``h`` followed along x by a block of cells of size ``h / 2``; every coarse face on
the interface plane meets 2 x 2 fine faces. Vertices that coincide are shared, so the topology builder of ``hexspace`` numbers the
fine edges and faces in the plane as entities of their own: their dofs are the SLAVES. A slave dof takes the value its dof
functional gives on the master side's field. For the interpolatory tensor basis of ``ND_HexahedronElement`` the functional of the
local dof (component c, node xi) is the c-th covariant reference component at xi, so with u = J_F^-T u^_F = J_C^-T u^_C

    x_slave * sign_F  =  [ J_F^T J_C^-T  sum_m Phi_m(xi_C) sign_C,m x[gid_C,m] ]_c .

The master trace is a polynomial the fine trace space contains (same order, affine sub-face), so the constrained space is
H(curl)-conforming; ``tests/test_nonconforming_cpu.py`` checks the tangential jump and the cavity eigenvalues."""
from __future__ import annotations

import dataclasses

import numpy as np
import scipy.sparse as sp

from . import hexmesh as hm
from . import hexspace as hs


@dataclasses.dataclass
class HangingBox:
    mesh: hm.HexMesh
    x_interface: float
    coarse: np.ndarray     # [NE] bool: element belongs to the coarse block
    size: tuple            # bounding box (Lx, Ly, Lz)


def hanging_box_mesh(nc=(1, 1, 1), nfx=2, h=1.0, scramble_seed=None, n_attr=1) -> HangingBox:
    """``nc`` = (ncx, ny, nz) coarse cells of size h on x in [0, ncx h]; (nfx, 2 ny, 2 nz) cells of size h / 2 behind them."""
    ncx, ny, nz = nc
    cm = hm.box_mesh((ncx, ny, nz), size=(ncx * h, ny * h, nz * h))
    fm = hm.box_mesh((nfx, 2 * ny, 2 * nz), size=(nfx * h / 2, ny * h, nz * h), origin=(ncx * h, 0.0, 0.0))
    nvc = cm.verts.shape[0]
    cvid = lambda i, j, k: i + (ncx + 1) * (j + (ny + 1) * k)
    fvid = lambda i, j, k: i + (nfx + 1) * (j + (2 * ny + 1) * k)
    # fine vertices on the interface plane with even (j, k) ARE coarse vertices
    remap = nvc + np.arange(fm.verts.shape[0], dtype=np.int64)
    for k in range(0, 2 * nz + 1, 2):
        for j in range(0, 2 * ny + 1, 2):
            remap[fvid(0, j, k)] = cvid(ncx, j // 2, k // 2)
    used = np.unique(np.concatenate([cm.elems.ravel(), remap[fm.elems].ravel()]))
    compact = -np.ones(nvc + fm.verts.shape[0], dtype=np.int64)
    compact[used] = np.arange(used.size)
    verts = np.concatenate([cm.verts, fm.verts])[used]
    elems = compact[np.concatenate([cm.elems, remap[fm.elems]])]
    if scramble_seed is not None:  # random local frames (cube symmetries), as box_mesh does
        rng = np.random.default_rng(scramble_seed)
        choice = rng.integers(0, len(hm._ROTS), size=elems.shape[0])
        new = np.empty_like(elems)
        for e in range(elems.shape[0]):
            perm, flips = hm._ROTS[choice[e]]
            for c in range(2):
                for b in range(2):
                    for a in range(2):
                        newc = (a, b, c)
                        old = [0, 0, 0]
                        for d in range(3):
                            old[perm[d]] = 1 - newc[d] if flips[d] else newc[d]
                        new[e, a + 2 * b + 4 * c] = elems[e, old[0] + 2 * old[1] + 4 * old[2]]
        elems = new
    attr = (1 + (np.arange(elems.shape[0]) % n_attr)).astype(np.int32)
    coarse = np.arange(elems.shape[0]) < cm.ne
    mesh = hm.HexMesh(verts=verts, elems=elems, attr=attr)
    return HangingBox(mesh, ncx * h, coarse, (ncx * h + nfx * h / 2, ny * h, nz * h))


def _affine(mesh, e):
    """x = x0 + J xi for the straight-sided element e (columns of J: edge vectors from lattice corner 0)."""
    v = mesh.verts[mesh.elems[e]]
    J = np.stack([v[1] - v[0], v[2] - v[0], v[4] - v[0]], axis=1)
    return v[0], J


def _nd_basis_at(p, xi):
    """Phi[m, c]: reference vector value of lexicographic ND basis function m at the reference point xi (hexspace._nd_lex_layout)."""
    op, _ = hs.gauss_legendre(p)
    cp = hs.gauss_lobatto(p + 1)
    Bo = [hs.lagrange_table(op, [xi[d]])[0][0] for d in range(3)]
    Bc = [hs.lagrange_table(cp, [xi[d]])[0][0] for d in range(3)]
    lay = hs._nd_lex_layout(p)
    Phi = np.zeros((len(lay), 3))
    for m, (c, i, j, k) in enumerate(lay):
        ix = (i, j, k)
        v = 1.0
        for d in range(3):
            v *= Bo[d][ix[d]] if d == c else Bc[d][ix[d]]
        Phi[m, c] = v
    return Phi


@dataclasses.dataclass
class ConstrainedSpace:
    space: hs.HexSpace          # the L-vector space (masters and slaves); its ess_dofs are NOT meaningful here
    P: sp.csr_matrix            # [ndofs_L x n_true] conforming prolongation
    true_of: np.ndarray         # [ndofs_L] true-dof index or -1 for slaves
    ess_tdofs: np.ndarray       # essential TRUE dofs: everything on the boundary of the box
    slaves: np.ndarray          # L indices of the constrained dofs


def build_constrained_nd_space(hb: HangingBox, p: int, tol=1e-10) -> ConstrainedSpace:
    mesh = hb.mesh
    topo = hs.build_topology(mesh)
    space = hs.build_nd_space(mesh, topo, p)
    ne, P = space.lex_gid.shape
    lay = hs._nd_lex_layout(p)
    op, _ = hs.gauss_legendre(p)
    cp = hs.gauss_lobatto(p + 1)
    xI = hb.x_interface
    aff = [_affine(mesh, e) for e in range(ne)]
    ref_node = np.array([[op[ix[d]] if d == c else cp[ix[d]] for d in range(3)] for (c, *ix) in lay])  # [P, 3]
    # coarse elements touching the interface, with their bounding boxes (to locate the master of a point)
    masters = [e for e in range(ne) if hb.coarse[e] and np.any(np.abs(mesh.verts[mesh.elems[e], 0] - xI) < tol)]
    boxes = {e: (mesh.verts[mesh.elems[e]].min(axis=0), mesh.verts[mesh.elems[e]].max(axis=0)) for e in masters}
    rows = {}
    on_bdr = np.zeros(space.ndofs, dtype=bool)
    L = np.array(hb.size)
    for e in range(ne):
        x0, J = aff[e]
        for l in range(P):
            c = lay[l][0]
            xp = x0 + J @ ref_node[l]
            t = J[:, c]  # physical tangent direction of the dof
            # boundary of the box: node on a boundary plane and tangent inside that plane
            for d in range(3):
                if (abs(xp[d]) < tol or abs(xp[d] - L[d]) < tol) and abs(t[d]) < tol:
                    on_bdr[space.lex_gid[e, l]] = True
            if hb.coarse[e]:
                continue
            if abs(xp[0] - xI) > tol or abs(t[0]) > tol:
                continue  # not a tangential dof in the interface plane
            s = int(space.lex_gid[e, l])
            if s in rows:
                continue
            C = next(m for m in masters if np.all(xp >= boxes[m][0] - tol) and np.all(xp <= boxes[m][1] + tol))
            xc0, JC = aff[C]
            xiC = np.linalg.solve(JC, xp - xc0)
            T = J.T @ np.linalg.inv(JC).T            # u^_F = J_F^T J_C^-T u^_C
            w = _nd_basis_at(p, xiC) @ T[c]          # [P_master]
            w = w * space.lex_sign[C] * space.lex_sign[e, l]
            keep = np.abs(w) > 1e-14
            rows[s] = (space.lex_gid[C][keep], w[keep])
    slaves = np.array(sorted(rows), dtype=np.int64)
    is_slave = np.zeros(space.ndofs, dtype=bool)
    is_slave[slaves] = True
    true_of = -np.ones(space.ndofs, dtype=np.int64)
    true_of[~is_slave] = np.arange(int((~is_slave).sum()))
    n_true = int((~is_slave).sum())
    ri, ci, vi = [np.nonzero(~is_slave)[0]], [true_of[~is_slave]], [np.ones(n_true)]
    for s, (g, w) in rows.items():
        assert not is_slave[g].any(), "one level of hanging entities only: masters must be true dofs"
        ri.append(np.full(len(g), s))
        ci.append(true_of[g])
        vi.append(w)
    Pm = sp.coo_matrix((np.concatenate(vi), (np.concatenate(ri), np.concatenate(ci))), shape=(space.ndofs, n_true)).tocsr()
    Pm.sum_duplicates()
    ess = true_of[np.nonzero(on_bdr & ~is_slave)[0]]
    return ConstrainedSpace(space, Pm, true_of, np.sort(ess), slaves)


def build_constrained_h1_space(hb: HangingBox, p: int, tol=1e-10) -> ConstrainedSpace:
    """The auxiliary (H1) space of the same mesh: a fine-side node in the interface plane that is not a coarse node takes the value
    of the master's nodal interpolant there (continuity of the scalar potential; the discrete gradients of such functions are the
    constrained ND fields' gradients, so R_nd G P_h1 maps true dofs to true dofs)."""
    mesh = hb.mesh
    topo = hs.build_topology(mesh)
    space = hs.build_h1_space(mesh, topo, p)
    ne, P = space.lex_gid.shape
    n = p + 1
    cp = hs.gauss_lobatto(n)
    xI = hb.x_interface
    aff = [_affine(mesh, e) for e in range(ne)]
    ref_node = np.array([[cp[i], cp[j], cp[k]] for k in range(n) for j in range(n) for i in range(n)])
    masters = [e for e in range(ne) if hb.coarse[e] and np.any(np.abs(mesh.verts[mesh.elems[e], 0] - xI) < tol)]
    boxes = {e: (mesh.verts[mesh.elems[e]].min(axis=0), mesh.verts[mesh.elems[e]].max(axis=0)) for e in masters}
    coarse_gids = set(int(g) for e in range(ne) if hb.coarse[e] for g in space.lex_gid[e])
    rows = {}
    on_bdr = np.zeros(space.ndofs, dtype=bool)
    L = np.array(hb.size)
    for e in range(ne):
        x0, J = aff[e]
        for l in range(P):
            xp = x0 + J @ ref_node[l]
            if any(abs(xp[d]) < tol or abs(xp[d] - L[d]) < tol for d in range(3)):
                on_bdr[space.lex_gid[e, l]] = True
            s = int(space.lex_gid[e, l])
            if hb.coarse[e] or abs(xp[0] - xI) > tol or s in coarse_gids or s in rows:
                continue
            C = next(m for m in masters if np.all(xp >= boxes[m][0] - tol) and np.all(xp <= boxes[m][1] + tol))
            xc0, JC = aff[C]
            xiC = np.linalg.solve(JC, xp - xc0)
            B = [hs.lagrange_table(cp, [xiC[d]])[0][0] for d in range(3)]
            w = np.array([B[0][i] * B[1][j] * B[2][k] for k in range(n) for j in range(n) for i in range(n)])
            keep = np.abs(w) > 1e-14
            rows[s] = (space.lex_gid[C][keep], w[keep])
    slaves = np.array(sorted(rows), dtype=np.int64)
    is_slave = np.zeros(space.ndofs, dtype=bool)
    is_slave[slaves] = True
    n_true = int((~is_slave).sum())
    true_of = -np.ones(space.ndofs, dtype=np.int64)
    true_of[~is_slave] = np.arange(n_true)
    ri, ci, vi = [np.nonzero(~is_slave)[0]], [true_of[~is_slave]], [np.ones(n_true)]
    for s, (g, w) in rows.items():
        assert not is_slave[g].any()
        ri.append(np.full(len(g), s))
        ci.append(true_of[g])
        vi.append(w)
    Pm = sp.coo_matrix((np.concatenate(vi), (np.concatenate(ri), np.concatenate(ci))), shape=(space.ndofs, n_true)).tocsr()
    Pm.sum_duplicates()
    ess = true_of[np.nonzero(on_bdr & ~is_slave)[0]]
    return ConstrainedSpace(space, Pm, true_of, np.sort(ess), slaves)


def restriction_matrix(cs: ConstrainedSpace) -> sp.csr_matrix:
    """R [n_true x ndofs_L]: selects the true dofs of an L-vector (FiniteElementSpace::GetRestrictionMatrix on conforming dofs)."""
    tr = np.nonzero(cs.true_of >= 0)[0]
    return sp.csr_matrix((np.ones(tr.size), (cs.true_of[tr], tr)), shape=(cs.P.shape[1], cs.space.ndofs))
