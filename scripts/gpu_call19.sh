#!/bin/bash
# A/B: cp.async FIFO variants of sparse_grads (FREUD_SG_MODE=async) and decode_dacts (FREUD_DD_MODE=1|2)
set -u
O=gpurun_out/c19
mkdir -p $O
for m in 1 2; do
  FREUD_DD_MODE=$m timeout 600 python -m pytest tests/test_gpu_topk.py -m gpu -x -q -k "fused_decode_dacts or fast_path or bf16" > $O/pytest_dd$m.log 2>&1; tail -2 $O/pytest_dd$m.log
done
FREUD_SG_MODE=async timeout 600 python -m pytest tests/test_gpu_topk.py tests/test_gpu_bench_shapes.py -m gpu -x -q > $O/pytest_sg.log 2>&1; tail -2 $O/pytest_sg.log
run() {  # name, env...
  name=$1; shift
  for w in c3 c2; do
    env "$@" timeout 600 python bench.py --workload $w --no-cpu-baseline --no-eager --no-extras --profile-out $O/prof_${w}_$name.json > $O/bench_${w}_$name.json 2> $O/bench_${w}_$name.err
    python - <<PY
import json
try:
    txt=open("$O/bench_${w}_$name.json").read(); d=json.loads(txt[txt.index("{"):])
    p=json.load(open("$O/prof_${w}_$name.json"))["kernels"]
    print("$w $name", "ms/step", round(d["ms_per_step"],3), d["parity_check"]["ok"], d["clocks"]["sm_mhz"], d["clocks"]["reasons"], {k:round(p[k]["ms_per_step"],3) for k in ("freud_topk_decode_dacts","freud_topk_sparse_grads","freud_csc_build","freud_topk_encode")})
except Exception as ex:
    print("$w $name failed", ex); print(open("$O/bench_${w}_$name.err").read()[-1500:])
PY
  done
}
run base FREUD_X=0
run sgasync FREUD_SG_MODE=async
run dd1 FREUD_DD_MODE=1
run dd2 FREUD_DD_MODE=2
