#!/bin/bash
N=${1:-2}
cd /root/repo; mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 300 $TR tools/dist_check.py > gpurun_out/dist_check_${N}gpu.log 2>&1; tail -n 1 gpurun_out/dist_check_${N}gpu.log
B2P_HALO_TIMING=1 B2P_PDL=0 timeout 300 $TR bench.py --gpus $N --steps 200 --warmup 10 > gpurun_out/bench_${N}gpu_timing_nopdl.json 2> gpurun_out/bench_${N}gpu_timing_nopdl.err
grep "halo timing" gpurun_out/bench_${N}gpu_timing_nopdl.err | tail -2
timeout 300 $TR bench.py --gpus $N --steps 200 --warmup 10 > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err
python bench.py --steps 200 --warmup 10 --no-cpu-baseline --no-experiments > gpurun_out/bench_1gpu_same_box.json 2>> gpurun_out/bench_${N}gpu.err
for f in gpurun_out/bench_${N}gpu.json gpurun_out/bench_1gpu_same_box.json; do grep '^{' $f | cut -c1-230; done
