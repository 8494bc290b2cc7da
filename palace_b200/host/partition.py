"""Partition of a global hex space into per-rank L-vectors laid out [owned | ghosts grouped by owner]
plus the neighbour exchange lists of the shared-dof assembly (stand-in for mfem::ParFiniteElementSpace's
conforming prolongation, which Palace uses through ParOperator, /root/reference/palace/linalg/rap.cpp:212-222,
after mesh::Partition, /root/reference/palace/driver.cpp:66-70). A dof is owned by the lowest rank
whose elements touch it."""
from __future__ import annotations

import dataclasses

import numpy as np

from .hexspace import HexSpace


@dataclasses.dataclass
class LocalSpace:
    rank: int
    elems: np.ndarray          # global element ids of this rank
    n_true: int
    n_ghost: int
    local_to_global: np.ndarray  # [n_true + n_ghost]
    space: HexSpace            # local space (lex_gid in local numbering, ndofs = n_true + n_ghost)
    ess_tdofs: np.ndarray      # essential dofs among the owned ones (T-vector indices)
    ess_ldofs: np.ndarray      # essential dofs among owned + ghosts (L-vector indices)
    nbr: np.ndarray            # neighbour ranks (ascending)
    send_counts: np.ndarray
    send_idx: np.ndarray       # owned local indices, concatenated per neighbour, ascending global id
    recv_counts: np.ndarray
    n_interior: int = 0        # elems[:n_interior] touch no ghost dof

    @property
    def lsize(self):
        return self.n_true + self.n_ghost


def dof_owner(space: HexSpace, elem_rank: np.ndarray, nranks: int) -> np.ndarray:
    owner = np.full(space.ndofs, nranks, dtype=np.int64)
    np.minimum.at(owner, space.lex_gid.ravel(), np.repeat(elem_rank.astype(np.int64), space.P))
    return owner


def interface_order(mesh_elems: np.ndarray, elem_rank: np.ndarray, rank: int):
    """Element order of one rank shared by ALL its spaces: elements that touch no vertex of another
    rank's elements first (they can never touch a ghost dof of any space), interface elements last."""
    nv = int(mesh_elems.max()) + 1
    vmin = np.full(nv, np.iinfo(np.int64).max, dtype=np.int64)
    vmax = np.full(nv, -1, dtype=np.int64)
    er = np.repeat(elem_rank.astype(np.int64), mesh_elems.shape[1])
    np.minimum.at(vmin, mesh_elems.ravel(), er)
    np.maximum.at(vmax, mesh_elems.ravel(), er)
    shared_v = vmin != vmax
    elems = np.nonzero(elem_rank == rank)[0]
    iface = shared_v[mesh_elems[elems]].any(axis=1)
    return np.concatenate([elems[~iface], elems[iface]]), int((~iface).sum())


def partition_space(space: HexSpace, elem_rank: np.ndarray, rank: int, nranks: int, owner: np.ndarray | None = None,
                    order=None) -> LocalSpace:
    """``order`` = (elems, n_interior) from interface_order() to share one element order across spaces."""
    owner = dof_owner(space, elem_rank, nranks) if owner is None else owner
    elems = np.nonzero(elem_rank == rank)[0] if order is None else order[0]
    gids = np.unique(space.lex_gid[elems])
    own_mask = owner[gids] == rank
    owned = gids[own_mask]
    ghosts = gids[~own_mask]
    gorder = np.lexsort((ghosts, owner[ghosts]))  # by owner rank, then global id
    ghosts = ghosts[gorder]
    l2g = np.concatenate([owned, ghosts])
    g2l = np.full(space.ndofs, -1, dtype=np.int64)
    g2l[l2g] = np.arange(l2g.size)
    # interior elements (touching no ghost dof) first: they can run while the ghost exchange is in flight
    touches_ghost = (g2l[space.lex_gid[elems]] >= owned.size).any(axis=1)
    if order is None:
        elems = np.concatenate([elems[~touches_ghost], elems[touches_ghost]])
        n_interior = int((~touches_ghost).sum())
    else:
        n_interior = int(order[1])
        assert not touches_ghost[:n_interior].any()
    lex_local = g2l[space.lex_gid[elems]]
    ess_mask = np.zeros(space.ndofs, dtype=bool)
    ess_mask[space.ess_dofs] = True
    ess_l = np.nonzero(ess_mask[l2g])[0]
    ess_t = ess_l[ess_l < owned.size]
    local = HexSpace(space.kind, space.p, int(l2g.size), space.P, lex_local, space.lex_sign[elems], space.dof_map, ess_l,
                     np.bincount(lex_local.ravel(), minlength=l2g.size))
    # receive side: ghosts grouped by owner
    g_owner = owner[ghosts]
    recv_from = np.unique(g_owner)
    # send side: for every other rank s, my owned dofs that s touches
    send_lists = {}
    for s in range(nranks):
        if s == rank:
            continue
        es = np.nonzero(elem_rank == s)[0]
        if es.size == 0:
            continue
        touched = np.unique(space.lex_gid[es])
        mine = touched[owner[touched] == rank]
        if mine.size:
            send_lists[s] = g2l[mine]  # ascending global id == the receiver's ghost order for owner `rank`
    nbr = np.array(sorted(set(recv_from.tolist()) | set(send_lists.keys())), dtype=np.int32)
    send_counts = np.array([send_lists[s].size if s in send_lists else 0 for s in nbr], dtype=np.int64)
    recv_counts = np.array([(g_owner == s).sum() for s in nbr], dtype=np.int64)
    send_idx = np.concatenate([send_lists[s] for s in nbr if s in send_lists]) if send_lists else np.zeros(0, dtype=np.int64)
    return LocalSpace(rank, elems, int(owned.size), int(ghosts.size), l2g, local, ess_t, ess_l, nbr, send_counts,
                      send_idx.astype(np.int32), recv_counts, n_interior)


# ------------------------------------------------------------------------------------------------
# Tetrahedral spaces (BASELINE configs 3 / 4 run on partitioned tet meshes): same layout and exchange lists, built from
# the native element-dof array of a TetSpace / H1TetSpace.
# ------------------------------------------------------------------------------------------------


@dataclasses.dataclass
class LocalTetSpace:
    rank: int
    elems: np.ndarray            # global element ids of this rank, interior elements first
    n_true: int
    n_ghost: int
    local_to_global: np.ndarray
    idx: np.ndarray              # [ne_local][P] int32 local dof of every element dof (native order)
    curl_orient: np.ndarray      # [ne_local][P][3] int8 rows (None for H1)
    ess_tdofs: np.ndarray
    ess_ldofs: np.ndarray
    nbr: np.ndarray
    send_counts: np.ndarray
    send_idx: np.ndarray
    recv_counts: np.ndarray
    n_interior: int = 0

    @property
    def lsize(self):
        return self.n_true + self.n_ghost


def partition_tet_space(space, elem_rank: np.ndarray, rank: int, nranks: int) -> LocalTetSpace:
    """`space`: tetspace.TetSpace or H1TetSpace (fields idx, ndofs, ess_dofs, optionally curl_orient)."""
    gid = space.idx.astype(np.int64)
    P = gid.shape[1]
    owner = np.full(space.ndofs, nranks, dtype=np.int64)
    np.minimum.at(owner, gid.ravel(), np.repeat(elem_rank.astype(np.int64), P))
    elems = np.nonzero(elem_rank == rank)[0]
    gids = np.unique(gid[elems])
    own_mask = owner[gids] == rank
    owned, ghosts = gids[own_mask], gids[~own_mask]
    ghosts = ghosts[np.lexsort((ghosts, owner[ghosts]))]
    l2g = np.concatenate([owned, ghosts])
    g2l = np.full(space.ndofs, -1, dtype=np.int64)
    g2l[l2g] = np.arange(l2g.size)
    touches_ghost = (g2l[gid[elems]] >= owned.size).any(axis=1)
    elems = np.concatenate([elems[~touches_ghost], elems[touches_ghost]])
    ess_mask = np.zeros(space.ndofs, dtype=bool)
    ess_mask[space.ess_dofs] = True
    ess_l = np.nonzero(ess_mask[l2g])[0]
    g_owner = owner[ghosts]
    send_lists = {}
    for s in range(nranks):
        if s == rank:
            continue
        es = np.nonzero(elem_rank == s)[0]
        if es.size == 0:
            continue
        touched = np.unique(gid[es])
        mine = touched[owner[touched] == rank]
        if mine.size:
            send_lists[s] = g2l[mine]
    nbr = np.array(sorted(set(np.unique(g_owner).tolist()) | set(send_lists.keys())), dtype=np.int32)
    send_counts = np.array([send_lists[s].size if s in send_lists else 0 for s in nbr], dtype=np.int64)
    recv_counts = np.array([(g_owner == s).sum() for s in nbr], dtype=np.int64)
    send_idx = np.concatenate([send_lists[s] for s in nbr if s in send_lists]) if send_lists else np.zeros(0, dtype=np.int64)
    co = getattr(space, "curl_orient", None)
    return LocalTetSpace(rank, elems, int(owned.size), int(ghosts.size), l2g, g2l[gid[elems]].astype(np.int32),
                         None if co is None else np.ascontiguousarray(co[elems]), ess_l[ess_l < owned.size], ess_l, nbr, send_counts,
                         send_idx.astype(np.int32), recv_counts, int((~touches_ghost).sum()))


def partition_tets_by_slabs(mesh_elems: np.ndarray, verts: np.ndarray, nranks: int, axis: int = 0) -> np.ndarray:
    """Element -> rank by equal-count slabs of the centroid coordinate (stand-in for METIS, driver.cpp:66-70)."""
    c = verts[mesh_elems].mean(axis=1)[:, axis]
    order = np.argsort(c, kind="stable")
    rank = np.empty(mesh_elems.shape[0], dtype=np.int64)
    rank[order] = (np.arange(order.size) * nranks) // order.size
    return rank
