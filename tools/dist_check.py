#!/usr/bin/env python
"""Multi-GPU parity check (run under torchrun, one rank per GPU):
  * distributed ParOperator::Mult (NCCL shared-dof assembly) == single-partition oracle apply
  * global dot product (device reduction + ncclAllReduce) == NumPy
  * FGMRES + p-multigrid (Chebyshev / Hiptmair smoothing, coarse PCG) solves the same system as the
    single-partition sparse direct solve, i.e. the result is partition independent.
Prints one line 'DIST_CHECK OK ...' on rank 0, raises otherwise."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import pyoracle as O  # noqa: E402
from palace_b200 import capi  # noqa: E402
from palace_b200.host import assemble as asm  # noqa: E402
from palace_b200.host import hexmesh as hm  # noqa: E402
from palace_b200.host import hexspace as hs  # noqa: E402
from palace_b200.host import partition as pt  # noqa: E402
from tests import common  # noqa: E402


def main():
    rank, world, lrank = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lrank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lrank))
    uid = [capi.Ctx.nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    ctx = capi.Ctx(lrank, nccl_uid=uid[0], rank=rank, nranks=world)
    capi.set_stream(ctx)
    parts = {2: (2, 1, 1), 4: (2, 2, 1), 8: (2, 2, 2)}[world]
    n = (4, 4, 2)
    p = 3
    prob = common.make_problem(n=n, p=p, scramble=5, warp=0.03, n_attr=2)
    elem_rank = hm.partition_box(n, parts)
    orders = asm.p_sequence(p)
    blob = common.coefficient(O.CURLCURL_MASS, 2, "matrix", a_mass=1.0, a_curl=0.5)
    blob_h1 = common.coefficient(O.H1_DIFFUSION, 2, "matrix", a_mass=1.0)
    nd = {q: hs.build_nd_space(prob.mesh, prob.topo, q) for q in orders}
    h1 = {q: hs.build_h1_space(prob.mesh, prob.topo, q) for q in orders}
    order = pt.interface_order(prob.mesh.elems, elem_rank, rank)
    lnd = {q: pt.partition_space(nd[q], elem_rank, rank, world, order=order) for q in orders}
    lh1 = {q: pt.partition_space(h1[q], elem_rank, rank, world, order=order) for q in orders}
    elems = lnd[p].elems
    nB, nG = prob.node_tabs
    geom = capi.Geom.hex(ctx, prob.xe[elems], prob.mesh.attr[elems], prob.mesh_order, prob.q1d, nB, nG, prob.tabs.qw)

    def halo_of(ls):
        return capi.Halo(ctx, ls.n_true, ls.n_ghost, ls.nbr, ls.send_counts, ls.send_idx, ls.recv_counts)

    hnd = {q: halo_of(lnd[q]) for q in orders}
    hh1 = {q: halo_of(lh1[q]) for q in orders}
    if os.environ.get("B2P_HALO_P2P", "1") == "1":
        def gather(blob):
            out = [None] * world
            dist.all_gather_object(out, blob)
            return out
        for q in orders:
            hnd[q].enable_p2p(gather)
            hh1[q].enable_p2p(gather)

    def par_op(kind, ls, halo, blob_, fine=None):
        sp = ls.space
        t = hs.tables_1d(sp.p, prob.q1d)
        if kind == O.H1_DIFFUSION:
            args = (sp.p, sp.ndofs, sp.lex_gid.astype(np.int32), None, None, None, t.Bc, t.Gc)
        else:
            idx, ori = sp.native_restriction()
            args = (sp.p, sp.ndofs, idx, ori, sp.dof_map, t.Bo, t.Bc, t.Gc)
        if fine is None:
            op = capi.Op.create(ctx, geom, kind, *args, blob_)
        else:
            op = fine.coarsen(*args)
        op.set_essential(ls.ess_ldofs)  # owned + ghost copies
        A = capi.Operator.par(ctx, ls.n_true, ls.lsize, [op], None, ls.ess_tdofs, 1, halo)
        A.set_interior(ls.n_interior)
        A.local_op = op
        return A

    A, AG = {}, {}
    A[p] = par_op(O.CURLCURL_MASS, lnd[p], hnd[p], blob)
    AG[p] = par_op(O.H1_DIFFUSION, lh1[p], hh1[p], blob_h1)
    for q in orders[:-1]:
        A[q] = par_op(O.CURLCURL_MASS, lnd[q], hnd[q], blob, fine=A[p].local_op)
        AG[q] = par_op(O.H1_DIFFUSION, lh1[q], hh1[q], blob_h1, fine=AG[p].local_op)

    # ---- distributed Mult vs global oracle ----
    Ao = common.oracle_matrix(prob, O.CURLCURL_MASS, blob, space=nd[p])
    x = np.random.default_rng(0).standard_normal(nd[p].ndofs)
    own = lnd[p].local_to_global[: lnd[p].n_true]
    xd = torch.from_numpy(x[own]).cuda()
    yd = torch.empty_like(xd)
    A[p].mult(xd, yd)
    torch.cuda.synchronize()
    y_ref = Ao @ x
    err_apply = float(np.abs(yd.cpu().numpy() - y_ref[own]).max() / np.abs(y_ref).max())
    dot = capi.vec_dot(ctx, xd, yd)
    err_dot = abs(dot - float(x @ y_ref)) / abs(float(x @ y_ref))

    # ---- complex operator on the partitioned space: A + 0.3i A0 as a wrapper of two real ParOperators ----
    Ai = capi.Operator.par(ctx, lnd[p].n_true, lnd[p].lsize, [A[p].local_op], [0.3], lnd[p].ess_tdofs, 0, hnd[p])
    W = capi.ComplexOperator.wrap(ctx, A[p], Ai)
    A0 = common.oracle_matrix(prob, O.CURLCURL_MASS, blob, space=nd[p], eliminate=False).tolil()
    A0[nd[p].ess_dofs, :] = 0
    A0[:, nd[p].ess_dofs] = 0
    xi = np.random.default_rng(2).standard_normal(nd[p].ndofs)
    xid = torch.from_numpy(xi[own]).cuda()
    zr, zi = torch.empty_like(xd), torch.empty_like(xd)
    W.mult(xd, xid, zr, zi)
    torch.cuda.synchronize()
    z_ref = Ao @ (x + 1j * xi) + 0.3j * (A0.tocsr() @ (x + 1j * xi))
    err_cplx = float(np.abs(zr.cpu().numpy() + 1j * zi.cpu().numpy() - z_ref[own]).max() / np.abs(z_ref).max())

    # ---- FGMRES + GMG, distributed ----
    def interp(in_ls, in_h, out_ls, out_h, comps):
        it = capi.Interp(ctx, asm.space_dict(in_ls.space), asm.space_dict(out_ls.space), comps)
        return capi.Operator.interp(ctx, it, in_h, in_ls.n_true, out_h, out_ls.n_true)

    G = [interp(lh1[q], hh1[q], lnd[q], hnd[q], asm.gradient_comps(q)) for q in orders]
    P = [interp(lnd[a], hnd[a], lnd[b], hnd[b], asm.nd_prolongation_comps(a, b)) for a, b in zip(orders[:-1], orders[1:])]
    coarse = capi.Solver.krylov(ctx, capi.CG, rel_tol=1e-13, max_it=5000)
    cj = capi.Solver.jacobi(ctx)
    cj.set_operator(A[orders[0]])
    coarse.set_preconditioner(cj)
    coarse.set_operator(A[orders[0]])
    M = capi.Solver.gmg(ctx, coarse, P, G, cycle_it=1, smooth_it=1, cheby_order=6)
    M.gmg_set_operators([A[q] for q in orders], [AG[q] for q in orders])
    K = capi.Solver.krylov(ctx, capi.FGMRES, rel_tol=1e-10, max_it=60, max_dim=60)
    K.set_operator(A[p])
    K.set_preconditioner(M)
    b = np.random.default_rng(1).standard_normal(nd[p].ndofs)
    b[nd[p].ess_dofs] = 0.0
    bd = torch.from_numpy(b[own]).cuda()
    sol = torch.zeros_like(bd)
    K.mult(bd, sol)
    st = K.stats()
    import scipy.sparse.linalg as spla

    x_ref = spla.spsolve(Ao.tocsc(), b)
    err_solve = float(np.linalg.norm(sol.cpu().numpy() - x_ref[own]) / np.linalg.norm(x_ref))
    errs = torch.tensor([err_apply, err_dot, err_solve, err_cplx], dtype=torch.float64, device="cuda")
    dist.all_reduce(errs, op=dist.ReduceOp.MAX)
    errs = errs.cpu().numpy()
    if rank == 0:
        print(f"DIST_CHECK world={world} apply_err={errs[0]:.2e} dot_err={errs[1]:.2e} solve_err={errs[2]:.2e} "
              f"fgmres_its={st['its']} converged={st['converged']} complex_apply_err={errs[3]:.2e}", flush=True)
    assert errs[0] < 1e-12 and errs[1] < 1e-12 and errs[2] < 1e-8 and errs[3] < 1e-12 and st["converged"], errs
    if rank == 0:
        print("DIST_CHECK OK", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
