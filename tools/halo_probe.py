#!/usr/bin/env python
"""Diagnostic: time the pieces of the partitioned apply (NCCL exchanges, interior / interface kernels)."""
import os, sys, json
import numpy as np, torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from palace_b200 import capi
from palace_b200.host import hexmesh as hm, partition as pt
import bench

def main():
    rank, world, lrank = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lrank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lrank))
    uid = [capi.Ctx.nccl_unique_id() if rank == 0 else None]; dist.broadcast_object_list(uid, src=0)
    ctx = capi.Ctx(lrank, nccl_uid=uid[0], rank=rank, nranks=world); capi.set_stream(ctx)
    parts = {2: (2, 1, 1), 4: (2, 2, 1), 8: (2, 2, 2)}[world]
    n = 29; gn = tuple(n * p for p in parts)
    prob = bench.build_problem(gn, 3, 0.0, size=tuple(float(p) for p in parts))
    er = hm.partition_box(gn, parts)
    ls = pt.partition_space(prob["nd"], er, rank, world, order=pt.interface_order(prob["mesh"].elems, er, rank))
    halo = capi.Halo(ctx, ls.n_true, ls.n_ghost, ls.nbr, ls.send_counts, ls.send_idx, ls.recv_counts)
    lv = torch.zeros(ls.lsize, dtype=torch.float64, device="cuda")
    def timeit(f, it=50):
        for _ in range(5): f()
        torch.cuda.synchronize(); dist.barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(it): f()
        b.record(); torch.cuda.synchronize()
        return a.elapsed_time(b) / it * 1e3
    t_fwd = t_rev = t_both = 0.0
    nd = ls.space
    geom = capi.Geom.hex(ctx, prob["xe"][ls.elems], prob["mesh"].attr[ls.elems], 1, prob["q1d"], prob["nB"], prob["nG"], prob["tabs"].qw)
    idx, ori = nd.native_restriction(); t = prob["tabs"]
    op = capi.Op.create(ctx, geom, capi.CURLCURL_MASS, 3, ls.lsize, idx, ori, nd.dof_map, t.Bo, t.Bc, t.Gc, prob["blob"])
    def gather(blob):
        out = [None] * world; dist.all_gather_object(out, blob); return out
    if os.environ.get("B2P_HALO_P2P", "1") == "1": halo.enable_p2p(gather)
    A = capi.Operator.par(ctx, ls.n_true, ls.lsize, [op], None, None, 1, halo); A.set_interior(ls.n_interior)
    A0 = capi.Operator.par(ctx, ls.n_true, ls.lsize, [op], None, None, 1, halo)
    x = torch.rand(ls.n_true, dtype=torch.float64, device="cuda"); y = torch.zeros_like(x)
    xg = torch.zeros(max(1, ls.n_ghost), dtype=torch.float64, device="cuda"); yg = torch.zeros_like(xg)
    t_full = timeit(lambda: A.mult(x, y))
    t_noov = timeit(lambda: A0.mult(x, y))
    t_int = timeit(lambda: op.apply_add_split(1.0, x, xg, y, yg, ls.n_true, 0, ls.n_interior))
    t_ifc = timeit(lambda: op.apply_add_split(1.0, x, xg, y, yg, ls.n_true, ls.n_interior, -1))
    t_all = timeit(lambda: op.apply_add_split(1.0, x, xg, y, yg, ls.n_true, 0, -1))
    t_ms = timeit(lambda: y.zero_())
    xl = torch.rand(ls.lsize, dtype=torch.float64, device="cuda"); yl = torch.zeros_like(xl)
    t_nosplit = timeit(lambda: op.apply_add(xl, yl))
    print(f"rank {rank}: nosplit_kernel={t_nosplit:.1f} us", flush=True)
    print(f"rank {rank}: mult_overlap={t_full:.1f} mult_nooverlap={t_noov:.1f} interior={t_int:.1f} interface={t_ifc:.1f} all={t_all:.1f} memset={t_ms:.1f} us", flush=True)
    print(f"rank {rank}: n_true={ls.n_true} n_ghost={ls.n_ghost} n_interior={ls.n_interior}/{ls.elems.size} send={ls.send_counts.tolist()} "
          f"fwd={t_fwd:.1f}us rev={t_rev:.1f}us both={t_both:.1f}us", flush=True)
    dist.destroy_process_group()
main()
