#!/bin/bash
N=${1:-2}
cd /root/repo; mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 300 $TR bench.py --gpus $N --steps 200 --warmup 10 > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err
grep '^{' gpurun_out/bench_${N}gpu.json | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['n_gpus'], d['value'], d['ms_per_step'], d.get('partition_parity'), d['config']['partition'][:80])"
tail -n 3 gpurun_out/bench_${N}gpu.err | cut -c1-300
timeout 300 python -m pytest tests/test_solvers_gpu.py -m gpu -x -q -k "multi_gpu" 2>&1 | tail -2
