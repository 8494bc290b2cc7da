#!/bin/bash
# full ncu capture (with source) of the hot kernel under the bench command; the report stays on the box (58 MB), its
# CSV pages come back. usage: r2_ncu.sh <tag> [env assignments...]
cd /root/repo; mkdir -p gpurun_out
TAG=${1:-nd}; shift
env "$@" ncu --set full --clock-control none --import-source on -k regex:nd_hex_apply -s 5 -c 1 -o /tmp/$TAG -f \
    python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-experiments > gpurun_out/ncu_$TAG.log 2>&1
ncu -i /tmp/$TAG.ncu-rep --page details > gpurun_out/${TAG}_details.txt 2>&1
ncu -i /tmp/$TAG.ncu-rep --page raw --csv > gpurun_out/${TAG}_raw.csv 2>&1
ncu -i /tmp/$TAG.ncu-rep --page source --csv --print-source sass > gpurun_out/${TAG}_source_sass.csv 2>&1
ncu -i /tmp/$TAG.ncu-rep --page source --csv --print-source cuda > gpurun_out/${TAG}_source_cuda.csv 2>&1
ls -la gpurun_out
