#!/bin/bash
# A/B of the ND kernel layouts + parity tests
cd /root/repo
timeout 240 python -m pytest tests/test_apply_gpu.py -m gpu -x -q 2>&1 | tail -4
for k in 3 4; do
  for p in 3; do
    echo "kernel $k order $p"
    B2P_ND_KERNEL=$k timeout 120 python bench.py --steps 200 --warmup 10 --order $p --no-cpu-baseline 2>&1 | tail -1
  done
done
for p in 2 4; do
  echo "kernel 4 order $p"
  timeout 120 python bench.py --steps 100 --warmup 10 --order $p --n $((p==1?64:(p==2?40:23))) --no-cpu-baseline 2>&1 | tail -1
done
