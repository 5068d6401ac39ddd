#!/bin/bash
set -u
mkdir -p gpurun_out/c1
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/c1/gpu.txt 2>&1
nproc >> gpurun_out/c1/gpu.txt; free -g >> gpurun_out/c1/gpu.txt
python -m pytest tests -m gpu -q -rs --durations=15 > gpurun_out/c1/pytest.log 2>&1
tail -40 gpurun_out/c1/pytest.log
python bench.py --steps 10 --warmup 3 > gpurun_out/c1/bench_c3.json 2> gpurun_out/c1/bench_c3.err
tail -c 3000 gpurun_out/c1/bench_c3.json; tail -5 gpurun_out/c1/bench_c3.err
python bench.py --workload c2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/c1/bench_c2.json 2> gpurun_out/c1/bench_c2.err
tail -c 1500 gpurun_out/c1/bench_c2.json; tail -5 gpurun_out/c1/bench_c2.err
( time python bench.py --impl reference --steps 3 --warmup 1 ) > gpurun_out/c1/bench_ref.json 2> gpurun_out/c1/bench_ref.err
cat gpurun_out/c1/bench_ref.json; tail -5 gpurun_out/c1/bench_ref.err
