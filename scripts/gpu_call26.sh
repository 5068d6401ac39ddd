#!/bin/bash
# final-state validation: whole GPU suite, then the DEFAULT bench line (parity check, eager arm, extras) and c2
set -u
O=gpurun_out/c26
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1
tail -4 $O/pytest.log
timeout 900 python bench.py > $O/bench_c3.json 2> $O/bench_c3.err
timeout 600 python bench.py --workload c2 --no-cpu-baseline --no-eager --no-extras > $O/bench_c2.json 2> $O/bench_c2.err
python - <<PY
import json
for w in ("c3","c2"):
    try:
        txt=open("$O/bench_%s.json"%w).read(); d=json.loads(txt[txt.index("{"):])
        print(w, "ms/step", round(d["ms_per_step"],3), "value", round(d["value"]/1e6,2), "e2e", round(d["e2e"]["value"]/1e6,2), round(d["e2e"]["ms_per_step"],3), "e2e32", round(d["e2e_fp32_input"]["value"]/1e6,2), d["parity_check"]["ok"], d["roofline"]["frac"], d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
        if d.get("extra"): print(json.dumps(d["extra"])[:1800])
        if d.get("torch_eager_b200"): print("eager", d["torch_eager_b200"]["ms_per_step"], "cpu", d["cpu_baseline"])
    except Exception as ex:
        print(w, "failed", ex); print(open("$O/bench_%s.err"%w).read()[-1500:])
PY
