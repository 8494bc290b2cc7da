#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_tet_gpu.py -m gpu -x -q 2>&1 | tail -1
for nt in 0 4; do B2P_DENSE_NT=$nt timeout 200 python tools/tet_bench.py --order 6 --n 11 --steps 10 2>> gpurun_out/tet.err | cut -c1-330; done
for p in 4 5; do timeout 200 python tools/tet_bench.py --order $p --n 14 --steps 10 2>> gpurun_out/tet.err | cut -c1-330; done
