#!/bin/bash
# Multi-GPU call of a round (gpurun --gpus N -- 'bash tools/round2_multi_gpu.sh N'): parity of the partitioned hex and tet
# operators, the weak-scaling bench with the fused halo kernels + graph replay (default) and with the round-1 sequence.
N=${1:-2}
cd /root/repo; mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 300 $TR tools/dist_check.py > gpurun_out/dist_check_${N}gpu.log 2>&1
timeout 300 $TR tools/tet_dist_check.py > gpurun_out/tet_dist_check_${N}gpu.log 2>&1
timeout 300 $TR bench.py --gpus $N --steps 200 --warmup 10 > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err
B2P_HALO_FUSED=0 B2P_GRAPH_STREAM=0 timeout 300 $TR bench.py --gpus $N --steps 200 --warmup 10 > gpurun_out/bench_${N}gpu_round1_sequence.json 2>> gpurun_out/bench_${N}gpu.err
B2P_GRAPH_STREAM=0 timeout 300 $TR bench.py --gpus $N --steps 200 --warmup 10 > gpurun_out/bench_${N}gpu_fused_eager.json 2>> gpurun_out/bench_${N}gpu.err
python bench.py --steps 200 --warmup 10 --no-cpu-baseline --no-experiments > gpurun_out/bench_1gpu_same_box.json 2>> gpurun_out/bench_${N}gpu.err
tail -2 gpurun_out/*dist_check_${N}gpu*.log; for f in gpurun_out/bench_*gpu*.json; do echo $f; cut -c1-260 $f; echo; done
