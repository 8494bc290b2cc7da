#!/bin/bash
# single-GPU measurements of round 2 beyond the headline: tets at >= 1M dofs, BASELINE config 2 at size, the reference arm
cd /root/repo; mkdir -p gpurun_out
cat /sys/fs/cgroup/cpu.max 2>/dev/null > gpurun_out/cgroup_cpu.txt; cat /sys/fs/cgroup/cpu/cpu.cfs_quota_us /sys/fs/cgroup/cpu/cpu.cfs_period_us 2>/dev/null >> gpurun_out/cgroup_cpu.txt
python bench.py --impl reference --steps 10 --warmup 2 > gpurun_out/ref_arm.json 2> gpurun_out/ref_arm.err
for nt in 1 2 4; do B2P_DENSE_NT=$nt timeout 300 python tools/tet_bench.py --order 3 --n 21 --steps 20 >> gpurun_out/tet_p3_1M.jsonl 2>> gpurun_out/tet.err; done
timeout 300 python tools/tet_bench.py --order 6 --n 11 --steps 10 >> gpurun_out/tet_p6_1M.jsonl 2>> gpurun_out/tet.err
timeout 900 python tools/cylinder_bench.py --order 4 --refine 3 --nev 4 --tol 1e-8 > gpurun_out/cylinder_p4_refine3.json 2> gpurun_out/cylinder.err
cut -c1-600 gpurun_out/ref_arm.json; cat gpurun_out/cgroup_cpu.txt; cat gpurun_out/tet_p3_1M.jsonl gpurun_out/tet_p6_1M.jsonl; cut -c1-1500 gpurun_out/cylinder_p4_refine3.json; tail -3 gpurun_out/cylinder.err gpurun_out/tet.err
