#!/bin/bash
# multi-GPU call: multi-rank parity tests + data-parallel bench (fused peer-memory exchange vs NCCL all-reduce)
set -u
G=${1:-2}
O=gpurun_out/c7_$G
mkdir -p $O
if [ "$G" = "2" ]; then
  timeout 900 python -m pytest tests/test_gpu_dp.py tests/test_gpu_sharded.py tests/test_gpu_search.py -m gpu -q -rs > $O/pytest.log 2>&1
  tail -8 $O/pytest.log
fi
for mode in fused nccl; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29611 \
    bench.py --gpus $G --steps 20 --warmup 5 --dp-mode $mode > $O/bench_c3_$mode.json 2> $O/bench_c3_$mode.err
  python - <<PY
import json
try:
    txt=open("$O/bench_c3_$mode.json").read(); d=json.loads(txt[txt.index("{"):])
    print("$mode", "ms/step", round(d["ms_per_step"],3), "Mtok/s", round(d["value"]/1e6,2), "e2e", round(d["e2e"]["value"]/1e6,2), "e2e32", round(d["e2e_fp32_input"]["value"]/1e6,2), d["config"]["dp_exchange"], d["parity_check"], d["extra"])
    print({k:round(v*d["ms_per_step"],3) for k,v in d["kernel_shares"].items()})
except Exception as ex:
    print("$mode failed", ex); print(open("$O/bench_c3_$mode.err").read()[-2000:])
PY
done
