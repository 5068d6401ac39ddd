#!/bin/bash
# double-buffered decode + dacts: kernel test, then C3 with 4 narrow parts x 8 warps vs 2 wide parts x 4 warps; C2
set -u
O=gpurun_out/c18
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_topk.py tests/test_gpu_bench_shapes.py -m gpu -x -q > $O/pytest.log 2>&1
tail -5 $O/pytest.log
for parts in 4 2; do
  FREUD_DD_PARTS=$parts timeout 600 python bench.py --no-cpu-baseline --no-eager --no-extras --profile-out $O/prof_c3_$parts.json > $O/bench_c3_$parts.json 2> $O/bench_c3_$parts.err
  python - <<PY
import json
try:
    txt=open("$O/bench_c3_$parts.json").read(); d=json.loads(txt[txt.index("{"):])
    print("c3 parts=$parts", "ms/step", round(d["ms_per_step"],3), "Mtok/s", round(d["value"]/1e6,2), d["parity_check"]["ok"], d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
    p=json.load(open("$O/prof_c3_$parts.json")); print({k:round(v["ms_per_step"],3) for k,v in p["kernels"].items()})
except Exception as ex:
    print("failed", ex); print(open("$O/bench_c3_$parts.err").read()[-2000:])
PY
done
timeout 600 python bench.py --workload c2 --no-cpu-baseline --no-eager --no-extras --profile-out $O/prof_c2.json > $O/bench_c2.json 2> $O/bench_c2.err
python - <<PY
import json
p=json.load(open("$O/prof_c2.json")); print("c2", p["ms_per_step"], {k:round(v["ms_per_step"],3) for k,v in p["kernels"].items()})
PY
