#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
python bench.py --impl reference --steps 10 --warmup 2 > gpurun_out/ref_arm.json 2> gpurun_out/ref_arm.err
cut -c1-200 gpurun_out/ref_arm.json
for k in 2 1; do
  B2P_DENSE_KERNEL=$k ncu --set full --clock-control none --import-source on -k regex:dense_apply -s 3 -c 1 -o /tmp/dense$k -f python tools/tet_bench.py --order 3 --n 21 --steps 3 > gpurun_out/ncu_dense$k.log 2>&1
  ncu -i /tmp/dense$k.ncu-rep --page details > gpurun_out/dense${k}_p3_details.txt 2>&1
  ncu -i /tmp/dense$k.ncu-rep --page source --csv --print-source sass > gpurun_out/dense${k}_p3_source_sass.csv 2>&1
done
B2P_COARSE_ASSEMBLED=1 timeout 600 python tools/cylinder_bench.py --order 4 --refine 3 --nev 4 --tol 1e-8 --coarse-tol 1e-4 > gpurun_out/cylinder_p4_refine3_assembled_coarse.json 2> gpurun_out/cylinder.err
cut -c1-1200 gpurun_out/cylinder_p4_refine3_assembled_coarse.json
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log
