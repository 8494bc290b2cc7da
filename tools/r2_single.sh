#!/bin/bash
# single-GPU measurements of round 2 beyond the headline: tets at >= 1M dofs, BASELINE config 2 at size, the reference arm
cd /root/repo; mkdir -p gpurun_out
./tools/ub/zero_ubench > gpurun_out/zero_ubench.log 2>&1
python bench.py --impl reference --steps 10 --warmup 2 > gpurun_out/ref_arm.json 2> gpurun_out/ref_arm.err
for k in 2 1; do B2P_DENSE_KERNEL=$k timeout 300 python tools/tet_bench.py --order 3 --n 21 --steps 20 >> gpurun_out/tet_p3_1M.jsonl 2>> gpurun_out/tet.err; done
for k in 2 1; do B2P_DENSE_KERNEL=$k timeout 300 python tools/tet_bench.py --order 6 --n 11 --steps 10 >> gpurun_out/tet_p6_1M.jsonl 2>> gpurun_out/tet.err; done
B2P_TRACE_KERNEL=1 timeout 300 python -m pytest tests/test_tet_gpu.py tests/test_dense_gpu.py tests/test_palace_glue.py tests/test_solvers_gpu.py -m gpu -x -q 2>&1 | tail -3 > gpurun_out/pytest_new.log
timeout 900 python tools/cylinder_bench.py --order 4 --refine 3 --nev 4 --tol 1e-8 > gpurun_out/cylinder_p4_refine3.json 2> gpurun_out/cylinder.err
python bench.py --steps 100 --warmup 10 > gpurun_out/bench_n1_full.json 2> gpurun_out/bench_n1_full.err
cat gpurun_out/zero_ubench.log | head -40; cut -c1-700 gpurun_out/ref_arm.json; cat gpurun_out/tet_p3_1M.jsonl gpurun_out/tet_p6_1M.jsonl; cat gpurun_out/pytest_new.log; cut -c1-1800 gpurun_out/cylinder_p4_refine3.json; tail -n 3 gpurun_out/cylinder.err gpurun_out/tet.err | cut -c1-300; cut -c1-300 gpurun_out/bench_n1_full.json
bash tools/r2_ncu.sh nd6_final
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-experiments > /dev/null 2>&1
