#!/usr/bin/env python
"""ND tetrahedron apply throughput through the dense-basis (DMMA) operator: BASELINE configs 3/4 element
type. Usage on the GPU box:  python tools/tet_bench.py --order 3 --n 12   (6 n^3 tets)."""
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
    ap.add_argument("--n", type=int, default=12)
    ap.add_argument("--geom-order", type=int, default=1)
    ap.add_argument("--warp", type=float, default=0.0)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--terms", action="store_true",
                    help="also time a0 K + a2 M as a ParOperator over two dense terms (one fused dense operator against one apply per "
                         "term) and K - w^2 (1 - i tan d) M + i w C as a complex operator (two coefficient sums against 2-4 applies per term)")
    args = ap.parse_args()
    import torch

    from palace_b200 import capi
    from palace_b200.host import coeff as cf
    from palace_b200.host import tetspace as ts

    t0 = time.time()
    mesh = ts.box_tet_mesh(args.n, jitter=0.2, scramble_seed=1, warp_amp=args.warp)
    sp = ts.build_nd_tet_space(mesh, args.order)
    interp, curl, qpts, qw = ts.nd_tet_tables(args.order)
    qd = ts.geom_qdata(mesh.node_coords(args.geom_order), mesh.attr, args.geom_order, qpts, qw)
    t_host = time.time() - t0
    ctx = capi.Ctx(0)
    geom = capi.Geom.general(ctx, qd)
    blob = cf.coeff_ctx_pair(cf.coeff_ctx(a=1.0), cf.coeff_ctx(a=1.0))
    op = capi.Op.create_dense(ctx, geom, capi.CURLCURL_MASS, sp.ndofs, sp.idx, None, interp, curl, blob, curl_orient=sp.curl_orient)
    x = torch.rand(sp.ndofs, dtype=torch.float64, device="cuda")
    y = torch.empty_like(x)
    flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device="cuda")
    for _ in range(5):
        op.apply(x, y)
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    for a, b in ev:
        flush.zero_()
        a.record()
        op.apply(x, y)
        b.record()
    torch.cuda.synchronize()
    ms = float(np.mean([a.elapsed_time(b) for a, b in ev]))
    P, Q = sp.P, len(qw)
    flops = 2.0 * 2 * 6 * Q * P * mesh.ne  # two GEMMs [6Q x P] per element
    extra = {}
    if args.terms:
        try:
            extra = time_term_sums(args, capi, cf, ctx, geom, sp, interp, curl, flush)
        except Exception as exc:  # the single-operator numbers above stay valid
            extra = {"terms_failed": f"{type(exc).__name__}: {str(exc)[:160]}"}
    print(json.dumps({**extra, "workload": f"ND tet p={args.order} curl-curl+mass apply, dense-basis DMMA operator", "tets": mesh.ne, "dofs": int(sp.ndofs),
                      "P": P, "Q": Q, "ms_per_apply": ms, "MDoF_per_s": sp.ndofs / ms / 1e3, "TFLOP_per_s_fp64": flops / ms / 1e9,
                      "host_setup_s": t_host}))


def time_term_sums(args, capi, cf, ctx, geom, sp, interp, curl, flush):
    """Sums of dense terms (BuildParSumOperator / ComplexParOperator over K, M, C on tets): fused against term by term."""
    import torch

    n = sp.ndofs
    K = capi.Op.create_dense(ctx, geom, capi.CURLCURL, n, sp.idx, None, None, curl, cf.coeff_ctx(a=1.0), curl_orient=sp.curl_orient)
    M = capi.Op.create_dense(ctx, geom, capi.ND_MASS, n, sp.idx, None, interp, None, cf.coeff_ctx(a=1.0), curl_orient=sp.curl_orient)
    Cd = capi.Op.create_dense(ctx, geom, capi.ND_MASS, n, sp.idx, None, interp, None, cf.coeff_ctx(a=0.3), curl_orient=sp.curl_orient)
    x, xi = torch.rand(n, dtype=torch.float64, device="cuda"), torch.rand(n, dtype=torch.float64, device="cuda")
    y, yi = torch.empty_like(x), torch.empty_like(x)

    def timed(fn):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        for a, b in ev:
            flush.zero_()
            a.record()
            fn()
            b.record()
        torch.cuda.synchronize()
        return float(np.mean([a.elapsed_time(b) for a, b in ev]))

    out = {}
    ref = None
    for fused in ("1", "0"):
        os.environ["B2P_SUM_FUSED"] = fused
        A = capi.Operator.par(ctx, n, n, [K, M], [1.0, -2.0], None, diag_policy=1)
        out["real_sum_K_M_ms_" + ("fused" if fused == "1" else "per_term")] = timed(lambda: A.mult(x, y))
        out["real_sum_is_fused" if fused == "1" else "real_sum_per_term_is_fused"] = bool(A.is_fused())
        if ref is None:
            ref = y.clone()
        else:
            out["real_sum_fused_vs_per_term_rel_diff"] = float((y - ref).norm() / ref.norm())
    os.environ.pop("B2P_SUM_FUSED", None)
    ref = None
    for fused in ("1", "0"):
        os.environ["B2P_COMPLEX_FUSED"] = fused
        Z = capi.ComplexOperator.par(ctx, n, n, [K, M, Cd], [1.0, -2.0 + 0.1j, 0.7j], None, diag_policy=1)
        out["complex_sum_K_M_C_ms_" + ("two_sums" if fused == "1" else "per_term")] = timed(lambda: Z.mult(x, xi, y, yi))
        if ref is None:
            ref = (y.clone(), yi.clone())
        else:
            out["complex_sum_two_sums_vs_per_term_rel_diff"] = float(((y - ref[0]).norm() + (yi - ref[1]).norm()) / (ref[0].norm() + ref[1].norm()))
    os.environ.pop("B2P_COMPLEX_FUSED", None)
    return out


if __name__ == "__main__":
    main()
