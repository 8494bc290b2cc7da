#!/usr/bin/env python
"""Solver-loop benchmark on one GPU: FGMRES + p-multigrid (Chebyshev / Hiptmair smoothing, coarse PCG)
for (K + M) x = b on the uniform hex mesh; reports set-up time, V-cycle time, iterations, time to solution
and a per-kernel breakdown of one V-cycle (CUDA events)."""
import os
import argparse, json, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from palace_b200 import capi
from palace_b200.host import assemble as asm, coeff as cf, hexmesh as hm, hexspace as hs

def main():
    ap = argparse.ArgumentParser(); ap.add_argument("--n", type=int, default=29); ap.add_argument("--order", type=int, default=3)
    ap.add_argument("--cheby-order", type=int, default=0); ap.add_argument("--coarse-tol", type=float, default=1e-2)
    a = ap.parse_args()
    p = a.order; corder = a.cheby_order or max(2 * p, 4)
    ctx = capi.Ctx(0); capi.set_stream(ctx)
    t0 = time.time()
    mesh = hm.box_mesh(a.n, (1.0, 1.0, 1.0)); topo = hs.build_topology(mesh)
    orders = asm.p_sequence(p); q1d = p + 1
    nd = {q: hs.build_nd_space(mesh, topo, q) for q in orders}; h1 = {q: hs.build_h1_space(mesh, topo, q) for q in orders}
    nodes = hs.gauss_lobatto(2); xe = mesh.node_coords(1, nodes); qx, qw = hs.gauss_legendre(q1d); nB, nG = hs.lagrange_table(nodes, qx)
    t_host = time.time() - t0
    geom = capi.Geom.hex(ctx, xe, mesh.attr, 1, q1d, nB, nG, qw)
    blob = cf.coeff_ctx_pair(cf.coeff_ctx(a=1.0), cf.coeff_ctx(a=1.0)); blob_g = cf.coeff_ctx(a=1.0)
    def par(kind, sp, blob_, fine=None):
        t = hs.tables_1d(sp.p, q1d)
        if kind == capi.H1_DIFFUSION: args = (sp.p, sp.ndofs, sp.lex_gid.astype(np.int32), None, None, None, t.Bc, t.Gc)
        else:
            idx, ori = sp.native_restriction(); args = (sp.p, sp.ndofs, idx, ori, sp.dof_map, t.Bo, t.Bc, t.Gc)
        op = capi.Op.create(ctx, geom, kind, *args, blob_) if fine is None else fine.coarsen(*args)
        A = capi.Operator.par(ctx, sp.ndofs, sp.ndofs, [op], None, sp.ess_dofs, 1); A.local_op = op; return A
    torch.cuda.synchronize(); t0 = time.time()
    A = {p: par(capi.CURLCURL_MASS, nd[p], blob)}; AG = {p: par(capi.H1_DIFFUSION, h1[p], blob_g)}
    for q in orders[:-1]:
        A[q] = par(capi.CURLCURL_MASS, nd[q], blob, A[p].local_op); AG[q] = par(capi.H1_DIFFUSION, h1[q], blob_g, AG[p].local_op)
    G = [capi.Operator.interp(ctx, capi.Interp(ctx, asm.space_dict(h1[q]), asm.space_dict(nd[q]), asm.gradient_comps(q))) for q in orders]
    P = [capi.Operator.interp(ctx, capi.Interp(ctx, asm.space_dict(nd[x]), asm.space_dict(nd[y]), asm.nd_prolongation_comps(x, y))) for x, y in zip(orders[:-1], orders[1:])]
    torch.cuda.synchronize(); t_ops = time.time() - t0
    t0 = time.time()
    coarse = capi.Solver.krylov(ctx, capi.CG, rel_tol=a.coarse_tol, max_it=500); cj = capi.Solver.jacobi(ctx)
    coarse.set_check_interval(int(os.environ.get("B2P_COARSE_CG_CHECK", "1")))  # > 1: CG scalars stay on the device
    assembled = os.environ.get("B2P_COARSE_ASSEMBLED", "0") == "1"
    if assembled:  # PCG on the device-assembled p = 1 matrix (MfemWrapperSolver flow) instead of the matrix-free coarse operator
        coarse = capi.Solver.assembled(ctx, coarse, cj)
    else:
        cj.set_operator(A[orders[0]]); coarse.set_preconditioner(cj); coarse.set_operator(A[orders[0]])
    mg = capi.Solver.gmg(ctx, coarse, P, G, cycle_it=1, smooth_it=1, cheby_order=corder)
    mg.gmg_set_operators([A[q] for q in orders], [AG[q] for q in orders])
    torch.cuda.synchronize(); t_setup = time.time() - t0
    n = nd[p].ndofs
    b = torch.rand(n, dtype=torch.float64, device="cuda"); b[torch.from_numpy(nd[p].ess_dofs).cuda()] = 0.0
    x = torch.zeros_like(b); y = torch.zeros_like(b)
    def ev(f, it=5):
        f(); torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); [f() for _ in range(it)]; e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / it
    t_apply = ev(lambda: A[p].mult(b, y), 20)
    t_h1 = ev(lambda: AG[p].mult(torch.ones(h1[p].ndofs, dtype=torch.float64, device="cuda"), torch.zeros(h1[p].ndofs, dtype=torch.float64, device="cuda")), 10)
    xg = torch.rand(h1[p].ndofs, dtype=torch.float64, device="cuda")
    t_G = ev(lambda: G[-1].mult(xg, y), 10); t_Gt = ev(lambda: G[-1].mult_transpose(b, xg), 10)
    xc = torch.rand(nd[orders[-2]].ndofs, dtype=torch.float64, device="cuda")
    t_P = ev(lambda: P[-1].mult(xc, y), 10); t_Pt = ev(lambda: P[-1].mult_transpose(b, xc), 10)
    t_dot = ev(lambda: capi.vec_dot(ctx, b, y), 20); t_axpy = ev(lambda: capi.vec_axpby(ctx, 1.0, b, 1.0, y), 20)
    t_vc = ev(lambda: mg.mult(b, y), 3)
    K = capi.Solver.krylov(ctx, capi.FGMRES, rel_tol=1e-8, max_it=100, max_dim=100); K.set_operator(A[p]); K.set_preconditioner(mg)
    torch.cuda.synchronize(); t0 = time.time(); K.mult(b, x); torch.cuda.synchronize(); t_solve = time.time() - t0
    st = K.stats()
    A[p].mult(x, y); res = float(torch.linalg.norm(y - b) / torch.linalg.norm(b))
    print(json.dumps({"dofs": n, "orders": orders, "cheby_order": corder, "host_mesh_space_s": t_host, "operator_create_s": t_ops, "smoother_setup_s": t_setup,
                      "nd_apply_ms": t_apply, "h1_apply_ms": t_h1, "G_ms": t_G, "Gt_ms": t_Gt, "P_ms": t_P, "Pt_ms": t_Pt, "dot_ms": t_dot, "axpby_ms": t_axpy,
                      "vcycle_ms": t_vc, "coarse_level": ("assembled_csr nnz=%d" % mg.assembled_nnz()) if assembled else "matrix_free", "fgmres_its": st["its"], "solve_s": t_solve, "true_rel_residual": res, "converged": st["converged"]}))
main()
