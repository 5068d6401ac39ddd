#!/bin/bash
# N-GPU confirmation of the default bench line (fused data-parallel step, parity check, feature-sharded C4 extra)
set -u
G=${1:-8}
O=gpurun_out/c15_$G
mkdir -p $O
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29621 \
    bench.py --gpus $G --steps 20 --warmup 5 > $O/bench_c3.json 2> $O/bench_c3.err
python - <<PY
import json
try:
    txt=open("$O/bench_c3.json").read(); d=json.loads(txt[txt.index("{"):])
    print("N=$G ms/step", round(d["ms_per_step"],3), "Mtok/s", round(d["value"]/1e6,2), "e2e", round(d["e2e"]["value"]/1e6,2), "e2e32", round(d["e2e_fp32_input"]["value"]/1e6,2), d["config"]["dp_exchange"], d["parity_check"], d["extra"])
    print({k:round(v*d["ms_per_step"],3) for k,v in d["kernel_shares"].items()})
except Exception as ex:
    print("failed", ex); print(open("$O/bench_c3.err").read()[-3000:])
PY
