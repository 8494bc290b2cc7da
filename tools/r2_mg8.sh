#!/bin/bash
# 8-GPU call: parity of the partitioned hex / tet operators, weak-scaling bench lines (hex ND p=3; tets p=3 and p=6)
N=${1:-8}
cd /root/repo; mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 240 $TR tools/dist_check.py > gpurun_out/dist_check_${N}gpu.log 2>&1; tail -n 1 gpurun_out/dist_check_${N}gpu.log
timeout 240 $TR bench.py --gpus $N --steps 200 --warmup 10 > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err
grep '^{' gpurun_out/bench_${N}gpu.json | cut -c1-230
B2P_HALO_TIMING=1 B2P_PDL=0 timeout 240 $TR bench.py --gpus $N --steps 100 --warmup 10 > gpurun_out/bench_${N}gpu_timing_nopdl.json 2> gpurun_out/bench_${N}gpu_timing_nopdl.err
grep "halo timing" gpurun_out/bench_${N}gpu_timing_nopdl.err | sort | tail -8
timeout 240 $TR tools/tet_dist_check.py > gpurun_out/tet_dist_check_${N}gpu.log 2>&1; tail -n 1 gpurun_out/tet_dist_check_${N}gpu.log
timeout 400 $TR tools/tet_scale_bench.py --order 3 --n 16 --steps 20 > gpurun_out/tet_scale_p3_${N}gpu.json 2> gpurun_out/tet_scale_${N}gpu.err
timeout 400 $TR tools/tet_scale_bench.py --order 6 --n 8 --geom-order 2 --warp 0.03 --steps 10 > gpurun_out/tet_scale_p6_${N}gpu.json 2>> gpurun_out/tet_scale_${N}gpu.err
grep -h '^{' gpurun_out/tet_scale_p3_${N}gpu.json gpurun_out/tet_scale_p6_${N}gpu.json | cut -c1-600
tail -n 3 gpurun_out/tet_scale_${N}gpu.err | cut -c1-300
