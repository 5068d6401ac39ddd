#!/bin/bash
# epilogue-input prefetch (L2 prefetch warp + one-chunk-ahead registers), col_sum rewrite, interpolated masked top-k,
# rank-sorted CSC long lists: whole GPU suite, then AuxK-live / L1 per-call breakdowns and C2/C3 bench lines
set -u
O=gpurun_out/c43
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1
tail -4 $O/pytest.log
python scripts/aux_prof.py c3 > $O/aux_c3.log 2>&1; head -14 $O/aux_c3.log; tail -1 $O/aux_c3.log
python scripts/l1_prof.py > $O/l1.log 2>&1; head -12 $O/l1.log
for w in c3 c2; do
  timeout 600 python bench.py --workload $w --no-cpu-baseline --no-eager --no-extras --profile-out $O/prof_$w.json > $O/bench_$w.json 2> $O/bench_$w.err
  python - <<PY
import json
try:
    txt=open("$O/bench_$w.json").read(); d=json.loads(txt[txt.index("{"):])
    p=json.load(open("$O/prof_$w.json"))["kernels"]
    print("$w", "ms/step", round(d["ms_per_step"],3), d["parity_check"]["ok"], d["clocks"]["sm_mhz"], d["clocks"]["reasons"], {k:round(p[k]["ms_per_step"],3) for k in ("freud_topk_decode_dacts","freud_topk_sparse_grads","freud_csc_build","freud_topk_encode")})
except Exception as ex:
    print("$w failed", ex); print(open("$O/bench_$w.err").read()[-1500:])
PY
done
