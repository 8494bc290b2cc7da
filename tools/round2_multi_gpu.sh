#!/bin/bash
# Multi-GPU call of a round (gpurun --gpus N -- 'bash tools/round2_multi_gpu.sh N'): parity of the partitioned hex and tet
# operators, the weak-scaling bench with and without graph replay on an internal stream.
N=${1:-2}
cd /root/repo; mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 300 $TR tools/dist_check.py > gpurun_out/dist_check_${N}gpu.log 2>&1
timeout 300 $TR tools/tet_dist_check.py > gpurun_out/tet_dist_check_${N}gpu.log 2>&1
timeout 300 $TR bench.py --gpus $N --steps 200 --warmup 10 > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err
B2P_GRAPH_STREAM=1 timeout 300 $TR tools/dist_check.py > gpurun_out/dist_check_${N}gpu_graph.log 2>&1
B2P_GRAPH_STREAM=1 timeout 300 $TR bench.py --gpus $N --steps 200 --warmup 10 > gpurun_out/bench_${N}gpu_graph.json 2>> gpurun_out/bench_${N}gpu.err
tail -2 gpurun_out/*dist_check_${N}gpu*.log; cat gpurun_out/bench_${N}gpu.json gpurun_out/bench_${N}gpu_graph.json | cut -c1-300
