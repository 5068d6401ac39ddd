#!/bin/bash
# fused decode + dacts: kernel test, trainer parity tests, C3 / C2 bench lines, upload-clip entry points
set -u
O=gpurun_out/c13
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_topk.py tests/test_clip_helpers.py tests/test_gpu_bench_shapes.py -m gpu -x -q > $O/pytest.log 2>&1
tail -15 $O/pytest.log
timeout 600 python bench.py --no-cpu-baseline --no-eager --no-extras --profile-out $O/prof_c3.json > $O/bench_c3.json 2> $O/bench_c3.err
timeout 600 python bench.py --workload c2 --no-cpu-baseline --no-eager --no-extras --profile-out $O/prof_c2.json > $O/bench_c2.json 2> $O/bench_c2.err
python - <<PY
import json
for w in ("c3","c2"):
    try:
        txt=open("$O/bench_%s.json"%w).read(); d=json.loads(txt[txt.index("{"):])
        print(w, "ms/step", round(d["ms_per_step"],3), "Mtok/s", round(d["value"]/1e6,2), "e2e", round(d["e2e"]["value"]/1e6,2), d["roofline"]["frac"], d["parity_check"]["ok"], d["clocks"])
        p=json.load(open("$O/prof_%s.json"%w)); print({k:round(v["ms_per_step"],3) for k,v in p["kernels"].items()})
    except Exception as ex:
        print(w, "failed", ex); print(open("$O/bench_%s.err"%w).read()[-2000:])
PY
