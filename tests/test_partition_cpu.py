"""CPU: the N > 1 host logic. (1) a pure NumPy simulation of the [owned | ghost] exchange lists for
2, 4 and 8 partitions; (2) the same exchange over a real world_size-2 ``gloo`` process group. In both
the distributed apply (local oracle apply + P / P^T exchange) must reproduce the single-partition
operator -- the partition-independence the reference checks by running its regression cases at
several rank counts."""
import os
import sys

import numpy as np
import pytest

from oracle import pyoracle as O
from palace_b200.host import coeff as cf
from palace_b200.host import hexmesh as hm
from palace_b200.host import hexspace as hs
from palace_b200.host import partition as pt
from tests import common


def _setup(n=(4, 2, 2), p=2, parts=(2, 1, 1)):
    prob = common.make_problem(n=n, p=p, scramble=5, warp=0.03, n_attr=2)
    elem_rank = hm.partition_box(n, parts)
    blob = common.coefficient(O.CURLCURL_MASS, 2, "matrix")
    return prob, elem_rank, blob


def _local_apply(prob, ls, blob, x_l):
    interp, curl, _ = O.nd_hex_tables(prob.p, prob.q1d)
    idx, ori = ls.space.native_restriction()
    y = np.zeros(ls.lsize)
    qd = np.ascontiguousarray(prob.qdata_ref[ls.elems])
    return O.apply_add(O.CURLCURL_MASS, interp, curl, idx, ori, qd, blob, np.ascontiguousarray(x_l), y)


@pytest.mark.parametrize("parts", [(2, 1, 1), (2, 2, 1), (2, 2, 2)])
def test_partition_lists_reproduce_the_global_operator(parts):
    nranks = int(np.prod(parts))
    prob, elem_rank, blob = _setup(parts=parts)
    nd = prob.nd
    owner = pt.dof_owner(nd, elem_rank, nranks)
    L = [pt.partition_space(nd, elem_rank, r, nranks, owner) for r in range(nranks)]
    # every dof owned exactly once; ghosts grouped by ascending owner
    assert sum(l.n_true for l in L) == nd.ndofs
    for l in L:
        assert (np.diff(owner[l.local_to_global[l.n_true:]]) >= 0).all()
        assert np.array_equal(np.sort(l.nbr), l.nbr) and l.recv_counts.sum() == l.n_ghost
    x = np.random.default_rng(0).random(nd.ndofs)
    y_ref = common.oracle_apply(prob, O.CURLCURL_MASS, blob, x)
    # forward (P): owners send, ghosts receive in place
    xl = [np.concatenate([x[l.local_to_global[: l.n_true]], np.zeros(l.n_ghost)]) for l in L]
    for r, l in enumerate(L):
        off = 0
        for k, s in enumerate(l.nbr):
            rc = int(l.recv_counts[k])
            if rc:
                ls = L[s]
                ks = list(ls.nbr).index(r)
                so = int(ls.send_counts[:ks].sum())
                sent = xl[s][ls.send_idx[so: so + int(ls.send_counts[ks])]]
                assert sent.size == rc
                xl[r][l.n_true + off: l.n_true + off + rc] = sent
            off += rc
    for r, l in enumerate(L):
        assert np.allclose(xl[r], x[l.local_to_global])
    # local apply + reverse (P^T)
    yl = [_local_apply(prob, l, blob, xl[r]) for r, l in enumerate(L)]
    yo = [y[: l.n_true].copy() for y, l in zip(yl, L)]
    for r, l in enumerate(L):
        off = 0
        for k, s in enumerate(l.nbr):
            rc = int(l.recv_counts[k])
            if rc:
                ls = L[s]
                ks = list(ls.nbr).index(r)
                so = int(ls.send_counts[:ks].sum())
                np.add.at(yo[s], ls.send_idx[so: so + rc], yl[r][l.n_true + off: l.n_true + off + rc])
            off += rc
    for r, l in enumerate(L):
        assert np.abs(yo[r] - y_ref[l.local_to_global[: l.n_true]]).max() < 1e-12 * np.abs(y_ref).max()


def _gloo_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    prob, elem_rank, blob = _setup()
    nd = prob.nd
    l = pt.partition_space(nd, elem_rank, rank, world)
    x = np.random.default_rng(0).random(nd.ndofs)
    xl = torch.from_numpy(np.concatenate([x[l.local_to_global[: l.n_true]], np.zeros(l.n_ghost)]))

    def exchange(vec, reverse):
        reqs, bufs, off_s, off_r = [], [], 0, 0
        for k, s in enumerate(l.nbr):
            sc, rc = int(l.send_counts[k]), int(l.recv_counts[k])
            if not reverse:
                if sc:
                    reqs.append(dist.isend(vec[torch.from_numpy(l.send_idx[off_s: off_s + sc].astype(np.int64))].contiguous(), int(s)))
                if rc:
                    buf = torch.empty(rc, dtype=torch.float64)
                    bufs.append((buf, l.n_true + off_r, rc, None))
                    reqs.append(dist.irecv(buf, int(s)))
            else:
                if rc:
                    reqs.append(dist.isend(vec[l.n_true + off_r: l.n_true + off_r + rc].contiguous(), int(s)))
                if sc:
                    buf = torch.empty(sc, dtype=torch.float64)
                    bufs.append((buf, None, sc, l.send_idx[off_s: off_s + sc].astype(np.int64)))
                    reqs.append(dist.irecv(buf, int(s)))
            off_s += sc
            off_r += rc
        for r_ in reqs:
            r_.wait()
        for buf, pos, cnt, idx in bufs:
            if idx is None:
                vec[pos: pos + cnt] = buf
            else:
                vec.index_add_(0, torch.from_numpy(idx), buf)

    exchange(xl, reverse=False)
    yl = torch.from_numpy(_local_apply(prob, l, blob, xl.numpy()))
    exchange(yl, reverse=True)
    y_ref = common.oracle_apply(prob, O.CURLCURL_MASS, blob, x)
    err = float(np.abs(yl.numpy()[: l.n_true] - y_ref[l.local_to_global[: l.n_true]]).max() / np.abs(y_ref).max())
    # global dot product = sum over ranks of owned parts (vector.hpp:247-253)
    t = torch.tensor([float(yl.numpy()[: l.n_true] @ x[l.local_to_global[: l.n_true]])], dtype=torch.float64)
    dist.all_reduce(t)
    derr = abs(float(t) - float(y_ref @ x)) / abs(float(y_ref @ x))
    q.put((rank, err, derr))
    dist.destroy_process_group()


def test_gloo_world_size_2_distributed_apply():
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p_ in procs:
        p_.start()
    res = [q.get(timeout=180) for _ in range(2)]
    for p_ in procs:
        p_.join(timeout=60)
    for rank, err, derr in res:
        assert err < 1e-12 and derr < 1e-12, (rank, err, derr)


def test_p_sequence_matches_reference_coarsening():
    from palace_b200.host import assemble as asm

    # multigrid.hpp:44-70: LOGARITHMIC p -> (p + 1) / 2; SURVEY appendix: p=3 -> {1,2,3}; 4 -> {1,2,4}; 6 -> {1,2,3,6}
    assert asm.p_sequence(3) == [1, 2, 3]
    assert asm.p_sequence(4) == [1, 2, 4]
    assert asm.p_sequence(6) == [1, 2, 3, 6]
    assert asm.p_sequence(4, "linear") == [1, 2, 3, 4]


def test_tet_partition_reproduces_the_global_operator():
    """Partitioned ND tet space (3 ranks, slabs): every rank applies its local elements through the oracle with the local
    numbering + curl-oriented restriction, ghost contributions are summed into the owners (P^T) -- the owned parts must equal
    the global apply; the exchange lists are mutually consistent (what rank a sends to b is what b expects from a)."""
    import numpy as np

    from oracle import pyoracle as O
    from palace_b200.host import coeff as cf
    from palace_b200.host import partition as pt
    from palace_b200.host import tetspace as ts

    mesh = ts.box_tet_mesh((3, 2, 2), (1.0, 0.8, 0.9), jitter=0.2, scramble_seed=5)
    p = 2
    sp = ts.build_nd_tet_space(mesh, p)
    interp, curl, qpts, qw = ts.nd_tet_tables(p)
    qd = ts.geom_qdata(mesh.node_coords(1), mesh.attr, 1, qpts, qw)
    blob = cf.coeff_ctx_pair(cf.coeff_ctx(a=1.0), cf.coeff_ctx(a=1.0))
    x = np.random.default_rng(0).random(sp.ndofs)
    y_ref = O.apply_add_co(O.CURLCURL_MASS, interp, curl, sp.idx, sp.curl_orient, qd, blob, x, np.zeros(sp.ndofs))
    nr = 3
    er = pt.partition_tets_by_slabs(mesh.elems, mesh.verts, nr)
    loc = [pt.partition_tet_space(sp, er, r, nr) for r in range(nr)]
    assert sum(l.n_true for l in loc) == sp.ndofs
    yl = []
    for l in loc:
        xl = x[l.local_to_global]                                  # P: owners' values copied into the ghosts
        yl.append(O.apply_add_co(O.CURLCURL_MASS, interp, curl, l.idx, l.curl_orient, np.ascontiguousarray(qd[l.elems]), blob, xl,
                                 np.zeros(l.lsize)))
        assert not (l.idx[: l.n_interior] >= l.n_true).any()       # interior elements touch no ghost
    # P^T: ghost segments go back to their owners, in the order of the exchange lists
    y = np.zeros(sp.ndofs)
    for l, v in zip(loc, yl):
        y[l.local_to_global[: l.n_true]] += v[: l.n_true]
    for a in loc:
        off = a.n_true
        for k, b in enumerate(a.nbr):
            rc = int(a.recv_counts[k])
            ghost_gids = a.local_to_global[off:off + rc]
            lb = loc[int(b)]
            kb = list(lb.nbr).index(a.rank)
            so = int(lb.send_counts[:kb].sum())
            send_gids = lb.local_to_global[lb.send_idx[so:so + int(lb.send_counts[kb])]]
            assert (ghost_gids == send_gids).all()                # both sides agree on the segment and its order
            y[ghost_gids] += yl[a.rank][off:off + rc]
            off += rc
    assert np.linalg.norm(y - y_ref) < 1e-13 * np.linalg.norm(y_ref)
