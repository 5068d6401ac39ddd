#!/bin/bash
# 2-GPU call: multi-rank parity tests + data-parallel bench (fused peer-memory exchange vs NCCL all-reduce)
set -u
O=gpurun_out/c7
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_dp.py tests/test_gpu_sharded.py tests/test_gpu_search.py -m gpu -q -rs > $O/pytest.log 2>&1
tail -30 $O/pytest.log
for mode in fused nccl; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 \
    bench.py --gpus 2 --steps 20 --warmup 5 --dp-mode $mode > $O/bench_c3_2gpu_$mode.json 2> $O/bench_c3_2gpu_$mode.err
  tail -c 1800 $O/bench_c3_2gpu_$mode.json; tail -5 $O/bench_c3_2gpu_$mode.err
done
