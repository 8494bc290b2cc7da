#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
B="python bench.py --steps 100 --warmup 10 --no-cpu-baseline --no-experiments"
for cfg in 42fcx 42fce 43fcx 42fgx; do
  B2P_ND_KERNEL=6 B2P_ND6_CFG=$cfg B2P_TRACE_KERNEL=1 timeout 300 python -m pytest tests/test_apply_gpu.py -m gpu -x -q -k "not halfwarp" 2>&1 | tail -1 > gpurun_out/pytest_nd6_$cfg.log
  B2P_ND_KERNEL=6 B2P_ND6_CFG=$cfg $B > gpurun_out/nd6_$cfg.json 2> gpurun_out/nd6_$cfg.err
done
for f in gpurun_out/nd6_*.json; do echo -n "$f "; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(round(d["value"]), round(d["ms_per_step"]*1e3,2), round(d["roofline"]["kernel_ms"]*1e3,2), round(d["roofline"]["frac"],3))
except Exception as e: print("fail", e)
PY
done
cat gpurun_out/pytest_nd6_*.log
bash tools/r2_ncu.sh nd6_42fcx B2P_ND_KERNEL=6 B2P_ND6_CFG=42fcx
python bench.py --impl reference --steps 10 --warmup 2 > gpurun_out/ref_arm.json 2> gpurun_out/ref_arm.err
cut -c1-400 gpurun_out/ref_arm.json; nproc; lscpu | grep -E "Model name|Thread|Core|Socket"
