#!/bin/bash
# last GPU call of round 2: the high-order kernel (all kinds, split, launch shapes) and DivFreeSolver on hardware, component-wide
# barriers against CTA barriers, the cylinder case (p = 4 now runs nd_hex_apply7_kernel), one bench line at p = 6
cd /root/repo; mkdir -p gpurun_out
timeout 30 python -m pytest tests/test_apply_gpu.py tests/test_divfree_gpu.py -m gpu -x -q 2>&1 | tail -2
timeout 40 python tools/nd7_ab.py --reps 20 > gpurun_out/nd7_ab3.jsonl 2> gpurun_out/nd7_ab3.err
python - <<'PY'
import json
for l in open('gpurun_out/nd7_ab3.jsonl'):
    r=json.loads(l)
    print(r['order'], r['variant'], 'FAILED '+r['failed'] if 'failed' in r else '%.2f us  frac %.3f  diff %.1e'%(r['kernel_ms']*1e3, r['roofline_frac'], r['rel_diff_to_nd_hex_apply4']))
PY
tail -2 gpurun_out/nd7_ab3.err
timeout 40 python -m pytest tests/test_zcylinder_gpu.py -m gpu -x -q 2>&1 | tail -2
for p in 6 5; do
  n=$((p==6?15:18))
  timeout 25 python bench.py --order $p --n $n --steps 50 --no-experiments --no-cpu-baseline 2>> gpurun_out/bench_hi.err | tee -a gpurun_out/bench_high_orders.jsonl | python -c "
import sys, json
for l in sys.stdin:
    r = json.loads(l); print(r['config'].get('order'), r['value'], r['ms_per_step'], r['roofline']['kernel'], r['roofline']['kernel_ms'], r['roofline']['frac'])
"
done
tail -2 gpurun_out/bench_hi.err
