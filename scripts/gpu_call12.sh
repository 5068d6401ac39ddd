#!/bin/bash
set -u
O=gpurun_out/c12
mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q -rs > $O/pytest.log 2>&1; tail -12 $O/pytest.log
timeout 300 python scripts/aux_prof.py c3 > $O/aux_c3.log 2>&1; head -5 $O/aux_c3.log; tail -1 $O/aux_c3.log
timeout 300 python scripts/l1_prof.py > $O/l1.log 2>&1; head -12 $O/l1.log
timeout 400 python bench.py --steps 20 --warmup 5 > $O/bench_c3.json 2> $O/bench_c3.err; tail -c 2500 $O/bench_c3.json; tail -3 $O/bench_c3.err
