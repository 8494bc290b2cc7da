#!/bin/bash
# final single-GPU check of a round: the whole GPU suite, smoke, both bench arms
cd /root/repo; mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --impl reference --steps 10 --warmup 2 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; cut -c1-160 gpurun_out/bench_reference.json
( time python bench.py ) > gpurun_out/bench_default.jsonl 2> gpurun_out/bench_default.err; tail -n 4 gpurun_out/bench_default.err
cut -c1-330 gpurun_out/bench_default.jsonl
