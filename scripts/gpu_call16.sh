#!/bin/bash
# FFMA2 gather kernels: whole GPU suite, C3 / C2 bench lines, AuxK-live and L1 per-call breakdowns
set -u
O=gpurun_out/c16
mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -x -q -rs > $O/pytest.log 2>&1
tail -12 $O/pytest.log
for w in c3 c2; do
  timeout 600 python bench.py --workload $w --no-cpu-baseline --no-eager --no-extras --profile-out $O/prof_$w.json > $O/bench_$w.json 2> $O/bench_$w.err
  python - <<PY
import json
try:
    txt=open("$O/bench_$w.json").read(); d=json.loads(txt[txt.index("{"):])
    print("$w", "ms/step", round(d["ms_per_step"],3), "Mtok/s", round(d["value"]/1e6,2), "e2e", round(d["e2e"]["value"]/1e6,2), d["roofline"]["frac"], d["parity_check"]["ok"], d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
    p=json.load(open("$O/prof_$w.json")); print({k:round(v["ms_per_step"],3) for k,v in p["kernels"].items()})
except Exception as ex:
    print("$w failed", ex); print(open("$O/bench_$w.err").read()[-2000:])
PY
done
python scripts/aux_prof.py c3 > $O/aux_c3.log 2>&1; cat $O/aux_c3.log
python scripts/l1_prof.py > $O/l1.log 2>&1; cat $O/l1.log
