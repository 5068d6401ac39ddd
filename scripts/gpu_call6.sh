#!/bin/bash
set -u
O=gpurun_out/c6
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_topk.py tests/test_gpu_bench_shapes.py -m gpu -q -x -k "not search" > $O/pytest.log 2>&1
tail -4 $O/pytest.log
for shape in 75000,384,6144 48000,768,24576; do
  for v in 0 4 5; do
    ONLY_SHAPE=$shape FREUD_ENC_VARIANT=$v timeout 120 python scripts/enc_variants.py 2>&1 | tee -a $O/variants.log
  done
done
ONLY_SHAPE=75000,384,6144 timeout 300 ncu --set full --clock-control none --import-source on -k regex:sm100_topk_kernel -s 3 -c 1 -o $O/enc_c2_spec \
     python scripts/enc_variants.py > $O/ncu_c2.log 2>&1
ls -la $O
