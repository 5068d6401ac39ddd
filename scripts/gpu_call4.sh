#!/bin/bash
set -u
O=gpurun_out/c4
mkdir -p $O
python -m pytest tests/test_gpu_topk.py -m gpu -q -x -k "edge or tail or fast_path" > $O/pytest.log 2>&1
tail -3 $O/pytest.log
for shape in 75000,384,6144 48000,768,24576; do
  for v in 0 3; do
    for f in 0 30 2 4 8 16; do
      ONLY_SHAPE=$shape FREUD_ENC_VARIANT=$v FREUD_ENC_FLAGS=$f python scripts/enc_variants.py 2>&1 | sed "s/^/flags=$f /" | tee -a $O/variants.log
    done
  done
done
for v in 0 3; do
  ONLY_SHAPE=75000,384,6144 FREUD_ENC_VARIANT=$v ncu --set full --clock-control none --import-source on -k regex:sm100_gemm_kernel -s 3 -c 1 -o $O/enc_c2_v$v \
     python scripts/enc_variants.py > $O/ncu_c2_v$v.log 2>&1
done
ONLY_SHAPE=48000,768,24576 FREUD_ENC_VARIANT=3 ncu --set full --clock-control none --import-source on -k regex:sm100_gemm_kernel -s 3 -c 1 -o $O/enc_c3_v3 \
     python scripts/enc_variants.py > $O/ncu_c3_v3.log 2>&1
ls -la $O
