#!/bin/bash
# Round-2 A/B of the opt-in variants on one B200 (gpurun -- 'bash tools/r2_ab.sh'); one JSON line per variant in gpurun_out/.
cd /root/repo; mkdir -p gpurun_out
B="python bench.py --steps 100 --warmup 10 --no-cpu-baseline --no-experiments"
$B > gpurun_out/ab_default.json 2> gpurun_out/ab.err
B2P_ND_FWDCHAIN=1 $B > gpurun_out/ab_fwdchain.json 2>> gpurun_out/ab.err
B2P_PDL=1 $B > gpurun_out/ab_pdl.json 2>> gpurun_out/ab.err
B2P_ND_FWDCHAIN=1 B2P_PDL=1 $B > gpurun_out/ab_fwdchain_pdl.json 2>> gpurun_out/ab.err
B2P_ND_FWDCHAIN=1 timeout 300 python -m pytest tests/test_apply_gpu.py -m gpu -x -q 2>&1 | tail -3 > gpurun_out/pytest_fwdchain.log
$B --warp 0.05 > gpurun_out/ab_warp.json 2>> gpurun_out/ab.err
for p in 2 4; do $B --order $p --n $((p==2?44:22)) > gpurun_out/ab_p$p.json 2>> gpurun_out/ab.err; B2P_ND_FWDCHAIN=1 $B --order $p --n $((p==2?44:22)) > gpurun_out/ab_p${p}_fwd.json 2>> gpurun_out/ab.err; done
timeout 200 python tools/zfused_bench.py > gpurun_out/zfused_p3.json 2> gpurun_out/zfused.err
timeout 200 python tools/solver_bench.py > gpurun_out/solver_bench.json 2> gpurun_out/solver_bench.err
B2P_INTERP_OWNER=1 timeout 200 python tools/solver_bench.py > gpurun_out/solver_bench_owner.json 2>> gpurun_out/solver_bench.err
for f in gpurun_out/ab_*.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(d["value"], d["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline"]["frac"])
except Exception as e: print("fail", e)
PY
done
