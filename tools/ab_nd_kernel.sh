#!/bin/bash
# Parity tests, then A/B of ND apply kernel variants on the headline workload (run on the GPU box).
cd /root/repo
timeout 300 python -m pytest tests/test_apply_gpu.py -m gpu -x -q 2>&1 | tail -3
B2P_ND_GD=1 timeout 300 python -m pytest tests/test_apply_gpu.py -m gpu -x -q -k "matches_oracle or agree" 2>&1 | tail -2
for cfg in "4 0" "4 1" "5 0"; do
  set -- $cfg
  echo "kernel $1 gd $2"
  B2P_ND_KERNEL=$1 B2P_ND_GD=$2 timeout 120 python bench.py --steps 200 --warmup 10 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('  value %.0f MDoF/s  ms/step %.5f  kernel_ms %.5f  frac %.4f' % (d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac']))"
done
