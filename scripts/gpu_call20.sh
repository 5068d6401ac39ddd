#!/bin/bash
# CSC build changes (coalesced scan, queued long-list sort): tests + C3/C2 bench in both placements; ncu capture of the
# AuxK-live step's dense kernels and of the L1 step's kernels
set -u
O=gpurun_out/c20
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_topk.py tests/test_gpu_bench_shapes.py tests/test_gpu_feed_collect.py tests/test_gpu_dropin.py -m gpu -x -q > $O/pytest.log 2>&1
tail -4 $O/pytest.log
for mode in side serial; do
 for w in c3 c2; do
  FREUD_CSC_MODE=$mode timeout 600 python bench.py --workload $w --no-cpu-baseline --no-eager --no-extras --profile-out $O/prof_${w}_$mode.json > $O/bench_${w}_$mode.json 2> $O/bench_${w}_$mode.err
  python - <<PY
import json
try:
    txt=open("$O/bench_${w}_$mode.json").read(); d=json.loads(txt[txt.index("{"):])
    p=json.load(open("$O/prof_${w}_$mode.json"))["kernels"]
    print("$w $mode", "ms/step", round(d["ms_per_step"],3), d["parity_check"]["ok"], d["clocks"]["sm_mhz"], d["clocks"]["reasons"], {k:round(p[k]["ms_per_step"],3) for k in ("freud_topk_decode_dacts","freud_topk_sparse_grads","freud_csc_build","freud_topk_encode")})
except Exception as ex:
    print("$w $mode failed", ex); print(open("$O/bench_${w}_$mode.err").read()[-1500:])
PY
 done
done
ncu --set full --clock-control none -k regex:"sm100_gemm_kernel|row_topk_mask|col_sum|axpby|residual_kernel" -s 40 -c 12 -o $O/aux python scripts/aux_prof.py c3 > $O/ncu_aux.log 2>&1
ncu -i $O/aux.ncu-rep --page raw --csv > $O/aux.raw.csv 2>/dev/null; rm -f $O/aux.ncu-rep
ncu --set full --clock-control none -k regex:"sm100_gemm_kernel|split_operand|col_sum|l1_colnorm|sum_splits" -s 60 -c 14 -o $O/l1 python scripts/l1_prof.py > $O/ncu_l1.log 2>&1
ncu -i $O/l1.ncu-rep --page raw --csv > $O/l1.raw.csv 2>/dev/null; rm -f $O/l1.ncu-rep
ls -la $O
