#!/bin/bash
set -u
O=gpurun_out/c2
mkdir -p $O
python -m pytest tests/test_gpu_topk.py tests/test_gpu_bench_shapes.py -m gpu -q -x -k "not search" > $O/pytest.log 2>&1
tail -15 $O/pytest.log
for shape in 75000,384,6144 48000,768,24576 24000,1280,10240; do
  for v in "0 0" "0 1" "10 0" "3 0"; do
    set -- $v
    ONLY_SHAPE=$shape FREUD_ENC_VARIANT=$1 FREUD_ENC_FLAGS=$2 python scripts/enc_variants.py 2>&1 | sed "s/^/flags=$2 /" | tee -a $O/variants.log
  done
done
for spec in "c2:75000,384,6144" "c3:48000,768,24576"; do
  name=${spec%%:*}; shape=${spec##*:}
  ONLY_SHAPE=$shape ncu --set full --clock-control none --import-source on -k regex:sm100_gemm_kernel -s 3 -c 1 -o $O/enc_$name \
     python scripts/enc_variants.py > $O/ncu_$name.log 2>&1
  tail -2 $O/ncu_$name.log
done
ls -la $O
