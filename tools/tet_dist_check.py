#!/usr/bin/env python
"""Multi-GPU parity check of the TETRAHEDRAL path (run under torchrun, one rank per GPU; BASELINE configs 3 / 4 are partitioned
tet meshes): the distributed ParOperator::Mult of the dense-basis ND tet operator -- slab partition, owned | ghost L-vectors,
NVLink / NCCL shared-dof assembly -- must equal the single-partition oracle apply on the owned dofs of every rank.
  torchrun --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/tet_dist_check.py
Prints 'TET_DIST_CHECK OK ...' on rank 0, raises otherwise."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import pyoracle as O  # noqa: E402
from palace_b200 import capi  # noqa: E402
from palace_b200.host import coeff as cf  # noqa: E402
from palace_b200.host import partition as pt  # noqa: E402
from palace_b200.host import tetspace as ts  # noqa: E402


def main():
    rank, world, lrank = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lrank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lrank))
    uid = [capi.Ctx.nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    ctx = capi.Ctx(lrank, nccl_uid=uid[0], rank=rank, nranks=world)
    capi.set_stream(ctx)
    p = int(os.environ.get("TET_ORDER", "3"))
    mesh = ts.box_tet_mesh((2 * world, 3, 3), (float(world), 1.0, 1.0), jitter=0.2, scramble_seed=5)   # same on every rank
    sp = ts.build_nd_tet_space(mesh, p)
    interp, curl, qpts, qw = ts.nd_tet_tables(p)
    qd = ts.geom_qdata(mesh.node_coords(1), mesh.attr, 1, qpts, qw)
    blob = cf.coeff_ctx_pair(cf.coeff_ctx(a=1.0), cf.coeff_ctx(a=0.5))
    er = pt.partition_tets_by_slabs(mesh.elems, mesh.verts, world)
    ls = pt.partition_tet_space(sp, er, rank, world)
    halo = capi.Halo(ctx, ls.n_true, ls.n_ghost, ls.nbr, ls.send_counts, ls.send_idx, ls.recv_counts)
    if os.environ.get("B2P_HALO_P2P", "1") == "1":
        def gather(b):
            out = [None] * world
            dist.all_gather_object(out, b)
            return out
        halo.enable_p2p(gather)
    geom = capi.Geom.general(ctx, np.ascontiguousarray(qd[ls.elems]))
    op = capi.Op.create_dense(ctx, geom, O.CURLCURL_MASS, ls.lsize, ls.idx, None, interp, curl, blob, curl_orient=ls.curl_orient)
    op.set_essential(ls.ess_ldofs)
    A = capi.Operator.par(ctx, ls.n_true, ls.lsize, [op], None, ls.ess_tdofs, 1, halo)
    A.set_interior(ls.n_interior)
    # reference: global oracle apply with the essential rows / columns eliminated
    x = np.random.default_rng(0).random(sp.ndofs)
    xm = x.copy()
    xm[sp.ess_dofs] = 0.0
    y_ref = O.apply_add_co(O.CURLCURL_MASS, interp, curl, sp.idx, sp.curl_orient, qd, blob, xm, np.zeros(sp.ndofs))
    y_ref[sp.ess_dofs] = x[sp.ess_dofs]
    owned = ls.local_to_global[: ls.n_true]
    xd = torch.from_numpy(x[owned]).cuda()
    yd = torch.empty_like(xd)
    errs = []
    for _ in range(3):                                  # repeated: epochs / graphs of the exchange must stay consistent
        A.mult(xd, yd)
        torch.cuda.synchronize()
        errs.append(float(np.linalg.norm(yd.cpu().numpy() - y_ref[owned]) / np.linalg.norm(y_ref[owned])))
    e = torch.tensor([max(errs)], dtype=torch.float64, device="cuda")
    dist.all_reduce(e, op=dist.ReduceOp.MAX)
    assert float(e.item()) < 1e-12, f"rank {rank}: distributed tet apply differs from the oracle: {errs}"
    if rank == 0:
        print(f"TET_DIST_CHECK OK world={world} p={p} tets={mesh.ne} dofs={sp.ndofs} max rel err {float(e.item()):.2e}")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
