#!/bin/bash
# 8-GPU call: weak-scaling lines of the partitioned tet operator (BASELINE configs 3 / 4 element type and orders)
N=${1:-8}
cd /root/repo; mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 300 $TR tools/tet_scale_bench.py --order 3 --cells 16 --steps 20 > gpurun_out/tet_scale_p3_${N}gpu.json 2> gpurun_out/tet_scale_${N}gpu.err
timeout 300 $TR tools/tet_scale_bench.py --order 6 --cells 8 --geom-order 2 --warp 0.03 --steps 10 > gpurun_out/tet_scale_p6_${N}gpu.json 2>> gpurun_out/tet_scale_${N}gpu.err
grep -h '^{' gpurun_out/tet_scale_p3_${N}gpu.json gpurun_out/tet_scale_p6_${N}gpu.json | cut -c1-600
tail -n 3 gpurun_out/tet_scale_${N}gpu.err | cut -c1-300
