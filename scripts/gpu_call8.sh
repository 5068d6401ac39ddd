#!/bin/bash
set -u
O=gpurun_out/c8
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_topk.py tests/test_gpu_bench_shapes.py tests/test_gpu_l1.py -m gpu -q -x -k "not search" > $O/pytest.log 2>&1
tail -25 $O/pytest.log
timeout 300 python scripts/aux_prof.py c3 > $O/aux_c3.log 2>&1; head -30 $O/aux_c3.log
timeout 300 python scripts/aux_prof.py c2 > $O/aux_c2.log 2>&1; head -12 $O/aux_c2.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-eager --no-parity > $O/bench_c3.json 2> $O/bench_c3.err
python - <<'PY'
import json
for f in ("c3",):
    try:
        d=json.load(open(f"gpurun_out/c8/bench_{f}.json"))
        print(f, round(d["ms_per_step"],3), round(d["value"]/1e6,2), "e2e", round(d["e2e"]["value"]/1e6,2), round(d["roofline"]["frac"],3), d["clocks"]["reasons"])
    except Exception as ex:
        print(f, "failed", ex); print(open(f"gpurun_out/c8/bench_{f}.err").read()[-1500:])
PY
