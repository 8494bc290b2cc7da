#!/usr/bin/env python
"""Frequency sweep of a driven problem on tetrahedra in miniature (BASELINE config 3's workflow on one GPU): for every frequency
the system matrix  A(w) = K + i w C - w^2 M  and the real, positive preconditioner matrix  P(w) = K + w C + w^2 M  get new
coefficients IN PLACE (no operator is rebuilt: /root/reference/palace/drivers/drivensolver.cpp:176-198 re-assembles, here
b2p_coperator_set_coefficients / b2p_operator_par_set_coefficients refill the fused per-element tensors), the p-multigrid with
Hiptmair smoothing is handed the new level operators (SetOperators: diagonals + Chebyshev lambda_max estimates per level), and
complex FGMRES preconditioned by the real V-cycle on both parts (PCMatReal, spaceoperator.cpp:1098-1105) solves for one
right-hand side. K, M, C are dense-basis (DMMA) operators of ND tetrahedra of order p on a scrambled box mesh; C is a lossy mass.
Reports per frequency: re-coefficient + SetOperators time, solve time, iterations.

  python tools/driven_sweep_bench.py --order 3 --n 6 --freqs 4"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--order", type=int, default=3)
    ap.add_argument("--n", type=int, default=6, help="box of n^3 cells, 6 tets each")
    ap.add_argument("--freqs", type=int, default=4)
    ap.add_argument("--tol", type=float, default=1e-8)
    ap.add_argument("--max-it", type=int, default=100)
    args = ap.parse_args()
    import torch

    from palace_b200 import capi
    from palace_b200.host import assemble as asm
    from palace_b200.host import coeff as cf
    from palace_b200.host import tetspace as ts

    t0 = time.time()
    mesh = ts.box_tet_mesh(args.n, (1.0, 0.8, 0.9), jitter=0.2, scramble_seed=3)
    orders = asm.p_sequence(args.order)
    spaces = [ts.build_nd_tet_space(mesh, p) for p in orders]
    h1s = [ts.build_h1_tet_space(mesh, sp, sp.p) for sp in spaces]
    _, _, qpts, qw = ts.nd_tet_tables(orders[-1])  # every level integrates with the fine rule (shared q-data)
    qd = ts.geom_qdata(mesh.node_coords(1), mesh.attr, 1, qpts, qw)
    t_host = time.time() - t0
    ctx = capi.Ctx(0)
    capi.set_stream(ctx)
    geom = capi.Geom.general(ctx, qd)
    one, sig = cf.coeff_ctx(a=1.0), cf.coeff_ctx(a=0.05)
    Z, Ppar, auxs, grads, keep = None, [], [], [], []
    for li, (sp, h1) in enumerate(zip(spaces, h1s)):
        interp, curl = ts.nd_tet_element(sp.p).tabulate(qpts)
        mk = lambda kind, blob: capi.Op.create_dense(ctx, geom, kind, sp.ndofs, sp.idx, None, interp if kind != capi.CURLCURL else None,
                                                     curl if kind != capi.ND_MASS else None, blob, curl_orient=sp.curl_orient)
        K, Cd, M = mk(capi.CURLCURL, one), mk(capi.ND_MASS, sig), mk(capi.ND_MASS, one)
        keep += [K, Cd, M]
        Ppar.append(capi.Operator.par(ctx, sp.ndofs, sp.ndofs, [K, Cd, M], [1.0, 1.0, 1.0], sp.ess_dofs, diag_policy=1))
        if li == len(spaces) - 1:
            Z = capi.ComplexOperator.par(ctx, sp.ndofs, sp.ndofs, [K, Cd, M], [1.0, 1.0j, -1.0], sp.ess_dofs, diag_policy=1)
        # auxiliary space: G^T P G = H1 diffusion with the mass-type coefficients (w C + w^2 M); one term, rescaled per frequency
        _, grad = ts.h1_tet_element(h1.p).tabulate(qpts)
        aop = capi.Op.create_dense(ctx, geom, capi.H1_DIFFUSION, h1.ndofs, h1.idx, None, None, grad, one)
        keep.append(aop)
        auxs.append(capi.Operator.par(ctx, h1.ndofs, h1.ndofs, [aop], [1.0], h1.ess_dofs, diag_policy=1))
        git = capi.Interp.dense(ctx, ts.tet_discrete_gradient(sp.p), h1.idx, h1.ndofs, sp.idx, sp.ndofs, out_curl_orient=ts.dual_orient(sp))
        grads.append(capi.Operator.interp(ctx, git))
    prol = []
    for c, f in zip(spaces[:-1], spaces[1:]):
        it = capi.Interp.dense(ctx, ts.nd_tet_prolongation(c.p, f.p), c.idx, c.ndofs, f.idx, f.ndofs, in_curl_orient=c.curl_orient,
                               out_curl_orient=ts.dual_orient(f))
        prol.append(capi.Operator.interp(ctx, it))
    coarse = capi.Solver.krylov(ctx, capi.CG, rel_tol=1e-3, max_it=200)
    cj = capi.Solver.jacobi(ctx)
    coarse.set_preconditioner(cj)
    gmg = capi.Solver.gmg(ctx, coarse, prol, grads, cycle_it=1, smooth_it=1, cheby_order=max(2 * args.order, 4))
    pc = capi.ComplexSolver.real_pc(ctx, gmg)
    ksp = capi.ComplexSolver.krylov(ctx, capi.FGMRES, rel_tol=args.tol, max_it=args.max_it, max_dim=args.max_it)
    fine = spaces[-1]
    n = fine.ndofs
    rng = np.random.default_rng(0)
    b = rng.standard_normal(n)
    b[fine.ess_dofs] = 0.0
    br, bi = torch.from_numpy(b).cuda(), torch.zeros(n, dtype=torch.float64, device="cuda")
    xr, xi = torch.zeros_like(br), torch.zeros_like(br)
    rows = []
    for w in np.linspace(1.0, 3.0, args.freqs):  # below the first resonance of the box (w ~ 5)
        torch.cuda.synchronize()
        t1 = time.time()
        Z.set_coefficients([1.0, 1j * w, -w * w])
        for P_, A_ in zip(Ppar, auxs):
            P_.set_coefficients([1.0, w, w * w])
            A_.set_coefficients([0.05 * w + w * w])
        cj.set_operator(Ppar[0])           # (the coarse CG itself belongs to the multigrid now and gets its operator from it)
        gmg.gmg_set_operators(Ppar, auxs)
        ksp.set_operator(Z)
        ksp.set_preconditioner(pc)
        torch.cuda.synchronize()
        t2 = time.time()
        xr.zero_()
        xi.zero_()
        ksp.mult(br, bi, xr, xi)
        torch.cuda.synchronize()
        t3 = time.time()
        st = ksp.stats()
        rows.append({"omega": float(w), "set_operators_ms": 1e3 * (t2 - t1), "solve_ms": 1e3 * (t3 - t2), "its": st["its"],
                     "converged": st["converged"]})
    # residual of the last solve against the operator itself
    yr, yi = torch.empty_like(br), torch.empty_like(br)
    Z.mult(xr, xi, yr, yi)
    res = float(((yr - br).norm() ** 2 + (yi - bi).norm() ** 2).sqrt() / br.norm())
    print(json.dumps({"workload": f"driven sweep on ND tets p={args.order}: A(w) = K + i w C - w^2 M, complex FGMRES + real p-multigrid {orders} "
                                  f"(Chebyshev order {max(2 * args.order, 4)} + Hiptmair), coefficients updated in place per frequency",
                      "tets": int(mesh.ne), "complex_dofs": int(n), "frequencies": rows,
                      "mean_set_operators_ms": float(np.mean([r["set_operators_ms"] for r in rows])),
                      "mean_solve_ms": float(np.mean([r["solve_ms"] for r in rows])), "last_true_residual": res,
                      "complex_apply_fused_as_two_sums": Z.fused_applies() > 0, "host_setup_s": t_host}))


if __name__ == "__main__":
    main()
