#!/bin/bash
set -u
O=gpurun_out/c11
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_search.py tests/test_gpu_gemm.py tests/test_gpu_bench_shapes.py -m gpu -q -k "search or mask or c5" > $O/pytest.log 2>&1; tail -5 $O/pytest.log
timeout 120 python scripts/select_bench.py 2>&1 | tee $O/select.log
timeout 300 python scripts/aux_prof.py c3 > $O/aux_c3.log 2>&1; head -6 $O/aux_c3.log; tail -1 $O/aux_c3.log
timeout 600 python scripts/bench_extra.py > $O/bench_extra.json 2> $O/bench_extra.err; cat $O/bench_extra.json | head -120; tail -3 $O/bench_extra.err
