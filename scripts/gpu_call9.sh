#!/bin/bash
set -u
O=gpurun_out/c9
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_topk.py tests/test_gpu_bench_shapes.py tests/test_gpu_l1.py tests/test_gpu_feed_collect.py -m gpu -q -k "not search" > $O/pytest.log 2>&1
tail -30 $O/pytest.log
timeout 300 python scripts/aux_prof.py c3 > $O/aux_c3.log 2>&1; head -14 $O/aux_c3.log; tail -1 $O/aux_c3.log
timeout 300 python scripts/aux_prof.py c2 > $O/aux_c2.log 2>&1; tail -1 $O/aux_c2.log
timeout 300 python scripts/l1_prof.py > $O/l1.log 2>&1; cat $O/l1.log
