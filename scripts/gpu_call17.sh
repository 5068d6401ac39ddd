#!/bin/bash
# TMA-staged sparse weight gradients: parity tests + A/B bench against the register-staged kernel
set -u
O=gpurun_out/c17
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_topk.py tests/test_gpu_bench_shapes.py tests/test_gpu_feed_collect.py -m gpu -x -q > $O/pytest.log 2>&1
tail -5 $O/pytest.log
for mode in async regs; do
 for w in c3 c2; do
  FREUD_SG_MODE=$mode timeout 600 python bench.py --workload $w --no-cpu-baseline --no-eager --no-extras --profile-out $O/prof_${w}_$mode.json > $O/bench_${w}_$mode.json 2> $O/bench_${w}_$mode.err
  python - <<PY
import json
try:
    txt=open("$O/bench_${w}_$mode.json").read(); d=json.loads(txt[txt.index("{"):])
    print("$w $mode", "ms/step", round(d["ms_per_step"],3), "Mtok/s", round(d["value"]/1e6,2), "e2e", round(d["e2e"]["value"]/1e6,2), d["parity_check"]["ok"], d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
    p=json.load(open("$O/prof_${w}_$mode.json")); print({k:round(v["ms_per_step"],3) for k,v in p["kernels"].items()})
except Exception as ex:
    print("$w $mode failed", ex); print(open("$O/bench_${w}_$mode.err").read()[-2000:])
PY
 done
done
