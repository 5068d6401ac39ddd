#!/bin/bash
# CSC placement experiment + L2 probe + parity tests after the dataclass fix
set -u
O=gpurun_out/c14
mkdir -p $O
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/l2_probe scripts/l2_probe.cu && /tmp/l2_probe > $O/l2_probe.json; cat $O/l2_probe.json
timeout 900 python -m pytest tests/test_gpu_topk.py tests/test_clip_helpers.py tests/test_gpu_bench_shapes.py -m gpu -x -q > $O/pytest.log 2>&1
tail -8 $O/pytest.log
for mode in side serial side_hi; do
  FREUD_CSC_MODE=$mode timeout 600 python bench.py --no-cpu-baseline --no-eager --no-extras --profile-out $O/prof_c3_$mode.json > $O/bench_c3_$mode.json 2> $O/bench_c3_$mode.err
  python - <<PY
import json
try:
    txt=open("$O/bench_c3_$mode.json").read(); d=json.loads(txt[txt.index("{"):])
    print("$mode", "ms/step", round(d["ms_per_step"],3), "Mtok/s", round(d["value"]/1e6,2), "e2e", round(d["e2e"]["value"]/1e6,2), d["roofline"]["frac"], d["parity_check"]["ok"], d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
    p=json.load(open("$O/prof_c3_$mode.json")); print({k:round(v["ms_per_step"],3) for k,v in p["kernels"].items()})
except Exception as ex:
    print("$mode failed", ex); print(open("$O/bench_c3_$mode.err").read()[-2000:])
PY
done
