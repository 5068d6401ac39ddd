#!/bin/bash
set -u
O=gpurun_out/c5
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_topk.py tests/test_gpu_bench_shapes.py -m gpu -q -x -k "not search" > $O/pytest.log 2>&1
tail -8 $O/pytest.log
for shape in 1000,64,512 75000,384,6144 48000,768,24576 24000,1280,10240 24000,1280,81920; do
  for v in 0 3; do
    ONLY_SHAPE=$shape FREUD_ENC_VARIANT=$v timeout 120 python scripts/enc_variants.py 2>&1 | tee -a $O/variants.log
  done
done
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-eager --no-parity > $O/bench_c3.json 2> $O/bench_c3.err
timeout 300 python bench.py --workload c2 --steps 20 --warmup 5 --no-cpu-baseline --no-eager --no-parity > $O/bench_c2.json 2> $O/bench_c2.err
python - <<'PY'
import json
for f in ("c3","c2"):
    try:
        d=json.load(open(f"gpurun_out/c5/bench_{f}.json"))
        print(f, round(d["ms_per_step"],3), round(d["value"]/1e6,2), "e2e", round(d["e2e"]["value"]/1e6,2), round(d["roofline"]["frac"],3), d["clocks"]["reasons"])
        print({k:round(v*d["ms_per_step"],3) for k,v in d["kernel_shares"].items()})
    except Exception as ex:
        print(f, "failed", ex); print(open(f"gpurun_out/c5/bench_{f}.err").read()[-1500:])
PY
timeout 300 python scripts/aux_prof.py c3 > $O/aux_c3.log 2>&1; cat $O/aux_c3.log
timeout 300 python scripts/aux_prof.py c2 > $O/aux_c2.log 2>&1; cat $O/aux_c2.log
ONLY_SHAPE=75000,384,6144 timeout 300 ncu --set full --clock-control none --import-source on -k regex:sm100_topk_kernel -s 3 -c 1 -o $O/enc_c2_spec \
     python scripts/enc_variants.py > $O/ncu_c2.log 2>&1
ONLY_SHAPE=48000,768,24576 timeout 300 ncu --set full --clock-control none --import-source on -k regex:sm100_topk_kernel -s 3 -c 1 -o $O/enc_c3_spec \
     python scripts/enc_variants.py > $O/ncu_c3.log 2>&1
timeout 400 bash scripts/sanitize.sh memcheck > $O/sanitize.log 2>&1; tail -5 $O/sanitize.log
ls -la $O
