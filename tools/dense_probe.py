#!/usr/bin/env python
"""Times the dense-basis DMMA operator against the sum-factorised kernel on the same hex mesh (p = 3)."""
import json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from oracle import pyoracle as O
from palace_b200 import capi
from tests import common
ctx = capi.Ctx(0)
prob = common.make_problem(n=(20, 20, 20), p=3, mesh_order=1, warp=0.0, scramble=None, n_attr=1)
blob = common.coefficient(O.CURLCURL_MASS, 1, "const")
sp = prob.nd
geom = capi.Geom.general(ctx, prob.qdata_ref)
interp, curl, _ = O.nd_hex_tables(3, 4); idx, ori = sp.native_restriction()
dense = capi.Op.create_dense(ctx, geom, O.CURLCURL_MASS, sp.ndofs, idx, ori, interp, curl, blob)
fast = common.gpu_op(ctx, common.gpu_geom(ctx, prob), prob, O.CURLCURL_MASS, blob)
x = torch.rand(sp.ndofs, dtype=torch.float64, device="cuda"); y = torch.zeros_like(x)
def ev(f, it=20):
    for _ in range(3): f()
    torch.cuda.synchronize(); a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); [f() for _ in range(it)]; b.record(); torch.cuda.synchronize(); return a.elapsed_time(b) / it
td, tf = ev(lambda: dense.apply_add(x, y)), ev(lambda: fast.apply_add(x, y))
flops = 2.0 * 2 * 384 * 144 * prob.mesh.ne  # two GEMMs of [384 x 144] per element
print(json.dumps({"elements": prob.mesh.ne, "dofs": sp.ndofs, "dense_dmma_ms": td, "sum_factorised_ms": tf,
                  "dense_GDoF_s": sp.ndofs / td / 1e6, "sum_factorised_GDoF_s": sp.ndofs / tf / 1e6, "dense_TFLOP_s": flops / td / 1e9}))
