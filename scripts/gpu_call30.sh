#!/bin/bash
set -u
O=gpurun_out/c30
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_l1.py tests/test_gpu_bench_shapes.py tests/test_gpu_dp.py tests/test_gpu_feed_collect.py tests/test_gpu_dropin.py tests/test_clip_helpers.py -m gpu -x -q > $O/pytest.log 2>&1
tail -4 $O/pytest.log
python scripts/l1_ab.py 2>&1 | tail -6
python scripts/enqueue_time.py 2>&1 | tail -3
