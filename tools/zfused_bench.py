#!/usr/bin/env python
"""Complex matvec of the lossy system  A = K - w^2 (1 - i tan d) M  (+ a conductivity-type mass term) on split
real/imag vectors: term-by-term (four real applies per term, the reference's ComplexWrapperOperator) against the fused
single-pass element kernel. GPU box:  python tools/zfused_bench.py [--order 3 --n 29]."""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--order", type=int, default=3)
    ap.add_argument("--n", type=int, default=29)
    ap.add_argument("--steps", type=int, default=100)
    args = ap.parse_args()
    import torch

    import bench
    from palace_b200 import capi
    from palace_b200.host import coeff as cf

    prob = bench.build_problem(args.n, args.order, 0.0)
    ctx = capi.Ctx(0)
    capi.set_stream(ctx)
    nd, t, p, q1d = prob["nd"], prob["tabs"], prob["p"], prob["q1d"]
    geom = capi.Geom.hex(ctx, prob["xe"], prob["mesh"].attr, prob["mesh_order"], q1d, prob["nB"], prob["nG"], t.qw)
    idx, ori = nd.native_restriction()
    mk = lambda kind, a: capi.Op.create(ctx, geom, kind, p, nd.ndofs, idx, ori, nd.dof_map, t.Bo, t.Bc, t.Gc, cf.coeff_ctx(a=a))
    K, M, C = mk(capi.CURLCURL, 1.0), mk(capi.ND_MASS, 1.0), mk(capi.ND_MASS, 0.3)
    coefs = [1.0 + 0.0j, -9.0 * (1 - 0.05j), 0.7j]
    N = nd.ndofs
    xr, xi = torch.rand(N, dtype=torch.float64, device="cuda"), torch.rand(N, dtype=torch.float64, device="cuda")
    flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device="cuda")
    out = {"workload": f"complex matvec K + lossy M + conductivity term, ND hex p={p}, {N} complex dofs"}
    res = {}
    for name, env in (("term_by_term", "0"), ("fused", "1")):
        os.environ["B2P_COMPLEX_FUSED"] = env
        A = capi.ComplexOperator.par(ctx, N, N, [K, M, C], coefs, nd.ess_dofs, 1)
        yr, yi = torch.empty_like(xr), torch.empty_like(xr)
        for _ in range(5):
            A.mult(xr, xi, yr, yi)
        torch.cuda.synchronize()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        for a, b in ev:
            flush.zero_()
            a.record()
            A.mult(xr, xi, yr, yi)
            b.record()
        torch.cuda.synchronize()
        ms = float(np.mean([a.elapsed_time(b) for a, b in ev]))
        res[name] = (yr.cpu().numpy() + 1j * yi.cpu().numpy())
        out[name] = {"ms_per_matvec": ms, "complex_MDoF_per_s": N / ms / 1e3, "fused_applies": A.fused_applies()}
    # two right-hand sides in one pass (the building block of a two-vector PCMatReal V-cycle) vs two applies
    KM = capi.Op.create(ctx, geom, capi.CURLCURL_MASS, p, nd.ndofs, idx, ori, nd.dof_map, t.Bo, t.Bc, t.Gc,
                        cf.coeff_ctx_pair(cf.coeff_ctx(a=1.0), cf.coeff_ctx(a=1.0)))
    y0, y1 = torch.zeros_like(xr), torch.zeros_like(xr)
    for name, fn in (("two_applies", lambda: (KM.apply_add(xr, y0), KM.apply_add(xi, y1))), ("pair_apply", lambda: KM.apply_add_pair(1.0, xr, xi, y0, y1))):
        for _ in range(5):
            fn()
        torch.cuda.synchronize()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        for a, b in ev:
            flush.zero_()
            a.record()
            fn()
            b.record()
        torch.cuda.synchronize()
        out[name] = {"ms": float(np.mean([a.elapsed_time(b) for a, b in ev]))}
    out["pair_speedup"] = out["two_applies"]["ms"] / out["pair_apply"]["ms"]
    out["rel_diff"] = float(np.linalg.norm(res["fused"] - res["term_by_term"]) / np.linalg.norm(res["term_by_term"]))
    out["speedup"] = out["term_by_term"]["ms_per_matvec"] / out["fused"]["ms_per_matvec"]
    print(json.dumps(out))


if __name__ == "__main__":
    main()
