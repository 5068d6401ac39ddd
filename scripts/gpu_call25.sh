#!/bin/bash
# TMA input boxes of the residual / mask epilogues: L1, GEMM, AuxK tests; per-call breakdowns
set -u
O=gpurun_out/c25
mkdir -p $O
timeout 1500 python -m pytest tests/test_gpu_l1.py tests/test_gpu_gemm.py tests/test_gpu_topk.py tests/test_gpu_bench_shapes.py tests/test_gpu_dropin.py -m gpu -x -q > $O/pytest.log 2>&1
tail -4 $O/pytest.log
python scripts/l1_prof.py > $O/l1.log 2>&1; head -12 $O/l1.log
python scripts/aux_prof.py c3 > $O/aux_c3.log 2>&1; head -12 $O/aux_c3.log; tail -1 $O/aux_c3.log
FREUD_NO_TMA_IN=1 python scripts/l1_prof.py > $O/l1_noin.log 2>&1; head -6 $O/l1_noin.log
