#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_tet_gpu.py tests/test_dense_gpu.py tests/test_zcylinder_tet_gpu.py -m gpu -x -q 2>&1 | tail -2
for k in 2 1; do B2P_DENSE_KERNEL=$k timeout 300 python tools/tet_bench.py --order 3 --n 21 --steps 20 >> gpurun_out/tet_p3_1M_v2.jsonl 2>> gpurun_out/tet.err; done
B2P_DENSE_KERNEL=2 timeout 300 python tools/tet_bench.py --order 6 --n 11 --steps 10 >> gpurun_out/tet_p6_1M_v2.jsonl 2>> gpurun_out/tet.err
timeout 300 python tools/tet_scale_bench.py --order 3 --cells 16 --steps 20 > gpurun_out/tet_scale_p3_1gpu.json 2>> gpurun_out/tet.err
timeout 300 python tools/tet_scale_bench.py --order 6 --cells 8 --geom-order 2 --warp 0.03 --steps 10 > gpurun_out/tet_scale_p6_1gpu.json 2>> gpurun_out/tet.err
cat gpurun_out/tet_p3_1M_v2.jsonl gpurun_out/tet_p6_1M_v2.jsonl gpurun_out/tet_scale_p3_1gpu.json gpurun_out/tet_scale_p6_1gpu.json | cut -c1-420; tail -n 3 gpurun_out/tet.err | cut -c1-300
