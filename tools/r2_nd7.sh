#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
timeout 100 python -m pytest tests/test_apply_gpu.py -m gpu -x -q -k "cta_per_batch" 2>&1 | tail -2
timeout 200 python tools/nd7_ab.py --reps 20 > gpurun_out/nd7_ab2.jsonl 2> gpurun_out/nd7_ab2.err
python - <<'PY'
import json
for l in open('gpurun_out/nd7_ab2.jsonl'):
    r=json.loads(l)
    print(r['order'], r['variant'], 'FAILED '+r['failed'] if 'failed' in r else '%.2f us  frac %.3f  diff %.1e'%(r['kernel_ms']*1e3, r['roofline_frac'], r['rel_diff_to_nd_hex_apply4']))
PY
tail -3 gpurun_out/nd7_ab2.err
# full ncu capture of the p = 6 kernel: launch 3 = LDG variant (timed one), launch 7 = TMA-staged variant
timeout 240 ncu --set full --clock-control none --import-source on -k regex:nd_hex_apply7 -s 3 -c 5 -o /tmp/nd7 -f \
    python tools/nd7_ab.py --orders 6 --configs "6:123d,123g" --reps 1 --skip-round1 > gpurun_out/ncu_nd7.log 2>&1
for i in 0 4; do
  ncu -i /tmp/nd7.ncu-rep --launch-skip $i --launch-count 1 --page details > gpurun_out/nd7_p6_launch${i}_details.txt 2>&1
  ncu -i /tmp/nd7.ncu-rep --launch-skip $i --launch-count 1 --page source --csv --print-source sass > gpurun_out/nd7_p6_launch${i}_source_sass.csv 2>&1
done
ncu -i /tmp/nd7.ncu-rep --page raw --csv > gpurun_out/nd7_p6_raw.csv 2>&1
ls -la gpurun_out | grep nd7; tail -5 gpurun_out/ncu_nd7.log | cut -c1-300
