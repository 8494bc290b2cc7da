#!/usr/bin/env python
"""bench.py -- curl-curl + mass operator apply throughput (BASELINE.json metric: MDoF/s).

One "step" = one y = A x of the matrix-free ND curl-curl+mass operator on the L-vector of one
mesh partition (ceed::Operator::Mult semantics: zero-fill + apply-add), synthetic uniform hex
mesh, p = 3, ~2M dofs per GPU (BASELINE configs[1]). `value` times the device path with inputs
resident in HBM; `e2e` times the same call through the C ABI with HOST buffers (pinned host x ->
device, apply, device y -> host) every step. `--impl reference` times the reference's own
algorithm (dense non-tensor basis, the CPU oracle port) on the host cores.

Launch:  python bench.py [--gpus N --steps K --warmup W]   or, for N > 1,
         python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "curl-curl+mass operator apply throughput"
UNIT = "MDoF/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b2p", choices=["b2p", "reference"])
    ap.add_argument("--order", type=int, default=3)
    ap.add_argument("--n", type=int, default=29, help="elements per direction per GPU (29 -> 2.02M dofs at p=3)")
    ap.add_argument("--assemble-qdata", type=int, default=0)
    ap.add_argument("--warp", type=float, default=0.0)
    ap.add_argument("--cpu-sample-elems", type=int, default=0, help="elements in the CPU baseline sample (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def build_problem(n, p, warp, mesh_order=1, origin=(0.0, 0.0, 0.0)):
    from palace_b200.host import coeff as cf
    from palace_b200.host import hexmesh as hm
    from palace_b200.host import hexspace as hs

    mesh = hm.box_mesh(n, (1.0, 1.0, 1.0), warp_amp=warp, n_attr=1, origin=origin)
    topo = hs.build_topology(mesh)
    nd = hs.build_nd_space(mesh, topo, p)
    q1d = p + 1
    nodes = hs.gauss_lobatto(mesh_order + 1)
    xe = mesh.node_coords(mesh_order, nodes)
    qx, qw = hs.gauss_legendre(q1d)
    nB, nG = hs.lagrange_table(nodes, qx)
    tabs = hs.tables_1d(p, q1d)
    # one material: mu^-1 = I (curl part), eps = 1 (mass part)  
    blob = cf.coeff_ctx_pair(cf.coeff_ctx(a=1.0), cf.coeff_ctx(a=1.0))
    return dict(mesh=mesh, topo=topo, nd=nd, q1d=q1d, xe=xe, nB=nB, nG=nG, tabs=tabs, blob=blob, mesh_order=mesh_order, p=p)


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(index)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        rows = [r for (t, r) in self.rows if t0 - 0.05 <= t <= t1 + 0.15] or [r for (_, r) in self.rows]
        for r in rows:
            f = [x.strip() for x in r.split(",")]
            try:
                sm.append(float(f[0]))
                smax = float(f[1])
                for nm, v in zip(names, f[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm)}


def measured_peak_gbs():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def cpu_reference_apply(prob, nthreads, sample_elems, repeats=1):
    """Time the reference algorithm (dense [3Q x P] interp/curl per element + reference QFunction
    arithmetic, the oracle port) on the host cores over a bounded sample of elements."""
    from oracle import pyoracle as O

    nd = prob["nd"]
    p, q1d = prob["p"], prob["q1d"]
    ne = prob["mesh"].ne
    ns = ne if sample_elems <= 0 else min(ne, sample_elems)
    interp, curl, _ = O.nd_hex_tables(p, q1d)
    idx, ori = nd.native_restriction()
    qd = O.geom_hex_qdata(prob["xe"][:ns], prob["mesh"].attr[:ns], prob["mesh_order"], q1d)
    x = np.random.default_rng(1).random(nd.ndofs)
    y = np.zeros(nd.ndofs)
    idx_s, ori_s = np.ascontiguousarray(idx[:ns]), np.ascontiguousarray(ori[:ns])
    O.apply_add_mt(nthreads, O.CURLCURL_MASS, interp, curl, idx_s[: max(1, ns // 20)], ori_s[: max(1, ns // 20)], qd, prob["blob"], x, y)
    best = float("inf")
    for _ in range(repeats):
        y[:] = 0.0
        t0 = time.perf_counter()
        O.apply_add_mt(nthreads, O.CURLCURL_MASS, interp, curl, idx_s, ori_s, qd, prob["blob"], x, y)
        best = min(best, time.perf_counter() - t0)
    dofs_equiv = nd.ndofs * (ns / ne)
    return dofs_equiv / best / 1e6, ns, best


def run_reference(args, rank, world):
    if rank != 0:
        return
    prob = build_problem(args.n, args.order, args.warp)
    cores = os.cpu_count() or 1
    ne = prob["mesh"].ne
    # bound each step to a few seconds of CPU work
    sample = args.cpu_sample_elems or min(ne, max(cores * 64, 4096))
    for _ in range(max(0, min(args.warmup, 1))):
        cpu_reference_apply(prob, cores, min(sample, 512))
    vals, t_tot = [], 0.0
    steps = max(1, args.steps)
    for _ in range(steps):
        v, ns, dt = cpu_reference_apply(prob, cores, sample)
        vals.append(v)
        t_tot += dt
    value = float(np.sum([prob["nd"].ndofs * (ns / ne)] * steps) / t_tot / 1e6)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t_tot / steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, prob, world),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{ns} of {ne} elements per step (dense non-tensor basis apply, oracle port of the libCEED /cpu/self path), dofs scaled by the element fraction"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def workload_config(args, prob, world):
    nd = prob["nd"]
    return {
        "workload": f"ND hex p={args.order} curl-curl+mass Mult on a uniform {args.n}^3-per-GPU box mesh, stored per-point geometry"
                    + (" (assembled D)" if args.assemble_qdata else " (J^-T, w detJ; coefficient applied on the fly)"),
        "order": args.order, "elements_per_gpu": int(prob["mesh"].ne), "dofs_per_gpu": int(nd.ndofs), "n_gpus": world,
        "vector": "L-vector", "l2_policy": "inputs larger than L2 (q-data + x + y + indices > 126 MB)"
        if prob["mesh"].ne * (10 * (args.order + 1) ** 3 * 8) > 126e6 else "L2 flushed between timed iterations",
    }


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist

    from palace_b200 import capi

    assert torch.cuda.is_available(), "bench.py needs a GPU (the product path has no CPU fallback)"
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    ctx = capi.Ctx(local_rank)

    prob = build_problem(args.n, args.order, args.warp)
    nd, p, q1d = prob["nd"], prob["p"], prob["q1d"]
    geom = capi.Geom.hex(ctx, prob["xe"], prob["mesh"].attr, prob["mesh_order"], q1d, prob["nB"], prob["nG"], prob["tabs"].qw)
    idx, ori = nd.native_restriction()
    t = prob["tabs"]
    op = capi.Op.create(ctx, geom, capi.CURLCURL_MASS, p, nd.ndofs, idx, ori, nd.dof_map, t.Bo, t.Bc, t.Gc, prob["blob"],
                        assemble=bool(args.assemble_qdata))
    N = nd.ndofs
    xh = torch.from_numpy(np.random.default_rng(1 + rank).random(N)).pin_memory()
    yh = torch.empty(N, dtype=torch.float64).pin_memory()
    xd = xh.cuda()
    yd = torch.empty_like(xd)
    stream = torch.cuda.current_stream()

    qbytes = prob["mesh"].ne * 10 * q1d ** 3 * 8
    flush = None
    if qbytes < 160e6:  # working set not safely larger than L2: flush explicitly
        flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device="cuda")

    def step_device():
        op.apply(xd, yd)

    def step_e2e():
        xd.copy_(xh, non_blocking=True)
        op.apply(xd, yd)
        yh.copy_(yd, non_blocking=True)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(3, args.warmup)):
        step_device()
    barrier()

    # ---- device-resident timing (value): CUDA events per step, L2 flushed between steps if needed ----
    sampler = ClockSampler(local_rank) if rank == 0 else None
    t_wall0 = time.time()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    for a, b in evs:
        if flush is not None:
            flush.zero_()
        a.record(stream)
        step_device()
        b.record(stream)
    barrier()
    t_wall1 = time.time()
    ms = np.array([a.elapsed_time(b) for a, b in evs])
    total_ms = float(ms.sum())
    clocks = sampler.stop(t_wall0, t_wall1) if sampler else None
    tt = torch.tensor([total_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    total_ms = float(tt.item())
    value = world * N * args.steps / (total_ms * 1e-3) / 1e6

    # ---- kernel-only timing for the roofline (memset excluded) ----
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(min(args.steps, 50))]
    for a, b in kev:
        yd.zero_()
        if flush is not None:
            flush.zero_()
        a.record(stream)
        op.apply_add(xd, yd)
        b.record(stream)
    torch.cuda.synchronize()
    k_ms = float(np.mean([a.elapsed_time(b) for a, b in kev]))
    abytes = op.algorithmic_bytes()
    peak, peak_src = measured_peak_gbs()
    achieved = abytes / (k_ms * 1e-3) / 1e9

    # ---- end to end through the C ABI with host buffers ----
    for _ in range(3):
        step_e2e()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2e_steps = min(args.steps, 50)
    e0.record(stream)
    for _ in range(e2e_steps):
        step_e2e()
    e1.record(stream)
    barrier()
    te = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = world * N * e2e_steps / (float(te.item()) * 1e-3) / 1e6

    line = None
    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": workload_config(args, prob, world),
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(N * 8), "d2h_bytes_per_step": int(N * 8)},
            "gpu_launches": int(args.steps),  # one nd_hex_apply kernel per step (+1 cudaMemset node, not ours)
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": None, "peak_source": peak_src, "kernel": "nd_hex_apply_kernel", "kernel_ms": k_ms,
                         "algorithmic_bytes_per_launch": int(abytes)},
        }
        if not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            sample = args.cpu_sample_elems or min(prob["mesh"].ne, max(cores * 64, 4096))
            v, ns, dt = cpu_reference_apply(prob, cores, sample)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": f"{ns} of {prob['mesh'].ne} elements, one dense-basis apply ({dt:.2f} s)"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
