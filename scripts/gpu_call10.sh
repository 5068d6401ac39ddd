#!/bin/bash
set -u
O=gpurun_out/c10
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_l1.py tests/test_gpu_bench_shapes.py -m gpu -q -k "l1" > $O/pytest.log 2>&1; tail -3 $O/pytest.log
timeout 300 python scripts/l1_prof.py > $O/l1.log 2>&1; head -14 $O/l1.log
timeout 120 python scripts/select_bench.py 2>&1 | tee $O/select.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:row_topk_mask_warp -s 2 -c 1 -o $O/select python scripts/select_bench.py > $O/ncu_select.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:sm100_gemm_kernel -s 6 -c 1 -o $O/l1dec python scripts/l1_prof.py > $O/ncu_l1.log 2>&1
ls -la $O
