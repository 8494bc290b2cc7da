#!/bin/bash
# First GPU call of a round (run under gpurun from the repo root): parity, headline bench, the experiments prepared
# on the CPU side, and the ncu evidence -- everything lands in gpurun_out/.
#   gpurun --timeout 1500 -- 'bash tools/round2_first_call.sh'
cd /root/repo; mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest_gpu.log
python bench.py --steps 200 --warmup 10 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_n1_reference.json 2>> gpurun_out/bench_n1.err
# prepared experiments: PDL overlap of the zero-fill, graph replay on an internal stream, fused complex matvec, tets
B2P_PDL=1 timeout 300 python -m pytest tests/test_apply_gpu.py tests/test_solvers_gpu.py -m gpu -x -q 2>&1 | tail -3 > gpurun_out/pytest_pdl.log
B2P_PDL=1 python bench.py --steps 200 --warmup 10 --no-cpu-baseline > gpurun_out/bench_n1_pdl.json 2> gpurun_out/bench_n1_pdl.err
B2P_ND_FWDCHAIN=1 timeout 300 python -m pytest tests/test_apply_gpu.py -m gpu -x -q 2>&1 | tail -3 > gpurun_out/pytest_fwdchain.log
B2P_ND_FWDCHAIN=1 python bench.py --steps 200 --warmup 10 --no-cpu-baseline > gpurun_out/bench_n1_fwdchain.json 2> gpurun_out/bench_n1_fwdchain.err
B2P_ND_FWDCHAIN=1 B2P_PDL=1 python bench.py --steps 200 --warmup 10 --no-cpu-baseline > gpurun_out/bench_n1_fwdchain_pdl.json 2>> gpurun_out/bench_n1_fwdchain.err
timeout 300 python tools/zfused_bench.py > gpurun_out/zfused_p3.json 2> gpurun_out/zfused.err
timeout 300 python tools/zfused_bench.py --order 1 --n 60 >> gpurun_out/zfused_p3.json 2>> gpurun_out/zfused.err
timeout 300 python tools/tet_bench.py --order 3 --n 14 > gpurun_out/tet_p3.json 2> gpurun_out/tet.err
timeout 300 python tools/tet_bench.py --order 6 --n 6 >> gpurun_out/tet_p3.json 2>> gpurun_out/tet.err
for nt in 2 4; do B2P_DENSE_NT=$nt timeout 300 python tools/tet_bench.py --order 3 --n 14 >> gpurun_out/tet_p3.json 2>> gpurun_out/tet.err; done
B2P_DENSE_NT=4 timeout 300 python -m pytest tests/test_dense_gpu.py tests/test_tet_gpu.py -m gpu -x -q 2>&1 | tail -2 > gpurun_out/pytest_dense_nt4.log
# solver loop: reference CG vs device-scalar CG on the coarse level
timeout 300 python tools/solver_bench.py > gpurun_out/solver_bench.json 2> gpurun_out/solver_bench.err
B2P_COARSE_CG_CHECK=8 timeout 300 python tools/solver_bench.py > gpurun_out/solver_bench_devcg.json 2>> gpurun_out/solver_bench.err
B2P_COARSE_ASSEMBLED=1 timeout 300 python tools/solver_bench.py > gpurun_out/solver_bench_assembled_coarse.json 2>> gpurun_out/solver_bench.err
# BASELINE configs 0 / 2 on the reference's cylinder mesh (level 0 is compared with the reference's stored eig.csv)
timeout 300 python tools/cylinder_bench.py --order 4 --refine 0 --nev 6 > gpurun_out/cylinder_p4_l0.json 2> gpurun_out/cylinder.err
timeout 600 python tools/cylinder_bench.py --order 4 --refine 2 --nev 4 --tol 1e-8 > gpurun_out/cylinder_p4_l2.json 2>> gpurun_out/cylinder.err
# ncu: launch list of the bench command, then one full capture of the apply kernel and of the fused complex kernel
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 10 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:nd_hex_apply4 -s 5 -c 1 -o gpurun_out/nd4_full -f \
    python bench.py --steps 10 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
B2P_COMPLEX_FUSED=1 ncu --set full --clock-control none --import-source on -k regex:nd_hex_apply4 -s 12 -c 1 -o gpurun_out/nd4z_full -f \
    python tools/zfused_bench.py --steps 3 > /dev/null 2>&1
ls -la gpurun_out
