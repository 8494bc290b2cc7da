#!/bin/bash
# nd6 (register-staged pipeline) on one B200: parity, launch-shape A/B, ncu of the default shape
cd /root/repo; mkdir -p gpurun_out
B2P_ND_KERNEL=6 B2P_TRACE_KERNEL=1 timeout 600 python -m pytest tests/test_apply_gpu.py -m gpu -x -q 2>&1 | tail -4 > gpurun_out/pytest_nd6.log
B="python bench.py --steps 100 --warmup 10 --no-cpu-baseline --no-experiments"
for cfg in 43f 43y 42f 42y 53f 62f 72f; do
  B2P_ND_KERNEL=6 B2P_ND6_CFG=$cfg $B > gpurun_out/nd6_$cfg.json 2> gpurun_out/nd6_$cfg.err
done
B2P_ND_KERNEL=6 B2P_PDL=1 $B > gpurun_out/nd6_43f_pdl.json 2> gpurun_out/nd6_pdl.err
for p in 2 4; do B2P_ND_KERNEL=6 $B --order $p --n $((p==2?44:22)) > gpurun_out/nd6_p$p.json 2>> gpurun_out/nd6_p.err; done
for f in gpurun_out/nd6_*.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(d["value"], d["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline"]["frac"])
except Exception as e: print("fail", e)
PY
done
bash tools/r2_ncu.sh nd6 B2P_ND_KERNEL=6
