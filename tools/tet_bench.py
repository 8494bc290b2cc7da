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
    print(json.dumps({"workload": f"ND tet p={args.order} curl-curl+mass apply, dense-basis DMMA operator", "tets": mesh.ne, "dofs": int(sp.ndofs),
                      "P": P, "Q": Q, "ms_per_apply": ms, "MDoF_per_s": sp.ndofs / ms / 1e3, "TFLOP_per_s_fp64": flops / ms / 1e9,
                      "host_setup_s": t_host}))


if __name__ == "__main__":
    main()
