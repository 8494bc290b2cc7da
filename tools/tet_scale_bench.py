#!/usr/bin/env python
"""Weak scaling of the ND TETRAHEDRON operator (BASELINE configs 3 / 4: partitioned tet meshes): ParOperator::Mult of the
dense-basis curl-curl+mass operator on a slab-partitioned box of tets, n^3 * 6 tets per GPU, shared dofs exchanged through
the peer-memory halo. One JSON line on rank 0: MDoF/s over all GPUs, TFLOP/s of the two element GEMMs, time per Mult
(max over ranks, CUDA events, L2 flushed between steps).
  python tools/tet_scale_bench.py --order 3 --cells 21                                  (1 GPU)
  torchrun --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/tet_scale_bench.py --order 6 --cells 11"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--order", type=int, default=3)
    ap.add_argument("--cells", dest="n", type=int, default=21, help="cells per direction per GPU (6 n^3 tets per GPU); (not --n: torchrun claims that prefix)")
    ap.add_argument("--geom-order", type=int, default=1, help="2: curved (quadratic) tets, as the spheres example")
    ap.add_argument("--warp", type=float, default=0.0)
    ap.add_argument("--steps", type=int, default=20)
    args = ap.parse_args()
    from palace_b200 import capi
    from palace_b200.host import coeff as cf
    from palace_b200.host import partition as pt
    from palace_b200.host import tetspace as ts

    rank, world, lrank = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(lrank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", lrank))
        uid = [capi.Ctx.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        ctx = capi.Ctx(lrank, nccl_uid=uid[0], rank=rank, nranks=world)
    else:
        ctx = capi.Ctx(lrank)
    capi.set_stream(ctx)
    p = args.order
    t0 = time.time()
    mesh = ts.box_tet_mesh((args.n * world, args.n, args.n), (float(world), 1.0, 1.0), jitter=0.2, scramble_seed=1, warp_amp=args.warp)
    sp = ts.build_nd_tet_space(mesh, p)
    interp, curl, qpts, qw = ts.nd_tet_tables(p)
    blob = cf.coeff_ctx_pair(cf.coeff_ctx(a=1.0), cf.coeff_ctx(a=1.0))
    if world > 1:
        er = pt.partition_tets_by_slabs(mesh.elems, mesh.verts, world)
        ls = pt.partition_tet_space(sp, er, rank, world)
        elems, n_true, lsize, idx, co, ess_l, ess_t = ls.elems, ls.n_true, ls.lsize, ls.idx, ls.curl_orient, ls.ess_ldofs, ls.ess_tdofs
        halo = capi.Halo(ctx, ls.n_true, ls.n_ghost, ls.nbr, ls.send_counts, ls.send_idx, ls.recv_counts)
        if os.environ.get("B2P_HALO_P2P", "1") == "1":
            def gather(b):
                out = [None] * world
                dist.all_gather_object(out, b)
                return out
            halo.enable_p2p(gather)
    else:
        elems, n_true, lsize, idx, co, ess_l, ess_t, halo = np.arange(mesh.ne), sp.ndofs, sp.ndofs, sp.idx, sp.curl_orient, sp.ess_dofs, sp.ess_dofs, None
    # q-data of this rank's elements only (the global array would be 8x the memory at 8 GPUs)
    coords = mesh.node_coords(args.geom_order)
    qd = ts.geom_qdata(np.ascontiguousarray(coords[elems]), mesh.attr[elems], args.geom_order, qpts, qw)
    t_host = time.time() - t0
    geom = capi.Geom.general(ctx, np.ascontiguousarray(qd))
    op = capi.Op.create_dense(ctx, geom, capi.CURLCURL_MASS, lsize, idx, None, interp, curl, blob, curl_orient=co)
    op.set_essential(ess_l)
    A = capi.Operator.par(ctx, n_true, lsize, [op], None, ess_t, 1, halo)
    if world > 1:
        A.set_interior(ls.n_interior)
    x = torch.rand(n_true, dtype=torch.float64, device="cuda")
    y = torch.empty_like(x)
    flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device="cuda")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(5):
        A.mult(x, y)
    barrier()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    for a, b in ev:
        flush.zero_()
        a.record()
        A.mult(x, y)
        b.record()
    barrier()
    tot = torch.tensor([sum(a.elapsed_time(b) for a, b in ev)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tot, op=dist.ReduceOp.MAX)
    ms = float(tot.item()) / args.steps
    P, Q = sp.P, len(qw)
    flops = 2.0 * 2 * 6 * Q * P * mesh.ne  # two GEMMs [6Q x P] per element, all ranks
    if rank == 0:
        print(json.dumps({"workload": f"ND tet p={p} curl-curl+mass ParOperator Mult, dense-basis DMMA operator, {6 * args.n ** 3} tets per GPU, "
                                      f"geometry order {args.geom_order}", "n_gpus": world, "tets": int(mesh.ne), "global_true_dofs": int(sp.ndofs),
                          "P": P, "Q": Q, "ms_per_mult": ms, "MDoF_per_s": sp.ndofs / ms / 1e3, "TFLOP_per_s_fp64": flops / ms / 1e9,
                          "TFLOP_per_s_fp64_per_gpu": flops / ms / 1e9 / world, "scaling": "weak", "host_setup_s": t_host}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
