#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
B2P_ND_KERNEL=6 timeout 600 python -m pytest tests/test_apply_gpu.py -m gpu -x -q 2>&1 | tail -4 > gpurun_out/pytest_nd6.log
B="python bench.py --steps 100 --warmup 10 --no-cpu-baseline --no-experiments"
for cfg in 42fgx 42fge 42ygx; do
  B2P_ND_KERNEL=6 B2P_ND6_CFG=$cfg $B > gpurun_out/nd6_$cfg.json 2> gpurun_out/nd6_$cfg.err
done
B2P_ND_KERNEL=6 B2P_PDL=1 $B > gpurun_out/nd6_def_pdl.json 2> gpurun_out/nd6_pdl.err
B2P_ND_KERNEL=6 $B --warp 0.05 > gpurun_out/nd6_def_warp.json 2>> gpurun_out/nd6_pdl.err
for p in 2 4; do B2P_ND_KERNEL=6 $B --order $p --n $((p==2?44:22)) > gpurun_out/nd6_p$p.json 2>> gpurun_out/nd6_p.err; done
for f in gpurun_out/nd6_*.json; do echo -n "$f "; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(round(d["value"]), round(d["ms_per_step"]*1e3,2), round(d["roofline"]["kernel_ms"]*1e3,2), round(d["roofline"]["frac"],3))
except Exception as e: print("fail", e)
PY
done
cat gpurun_out/pytest_nd6.log
bash tools/r2_ncu.sh nd6_42fgx B2P_ND_KERNEL=6
