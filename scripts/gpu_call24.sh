#!/bin/bash
# half-precision activations straight into the fused path: whole suite + default bench line (e2e is the number to watch)
set -u
O=gpurun_out/c24
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1
tail -4 $O/pytest.log
for w in c3 c2; do
  timeout 600 python bench.py --workload $w --no-cpu-baseline --no-eager --no-extras --profile-out $O/prof_$w.json > $O/bench_$w.json 2> $O/bench_$w.err
  python - <<PY
import json
try:
    txt=open("$O/bench_$w.json").read(); d=json.loads(txt[txt.index("{"):])
    p=json.load(open("$O/prof_$w.json"))["kernels"]
    print("$w", "ms/step", round(d["ms_per_step"],3), "value", round(d["value"]/1e6,2), "e2e", round(d["e2e"]["value"]/1e6,2), round(d["e2e"]["ms_per_step"],3), "e2e32", round(d["e2e_fp32_input"]["value"]/1e6,2), d["parity_check"]["ok"], d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
except Exception as ex:
    print("$w failed", ex); print(open("$O/bench_$w.err").read()[-1500:])
PY
done
