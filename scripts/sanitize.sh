#!/usr/bin/env bash
# compute-sanitizer pass over the kernels with hand-rolled synchronisation: the fused
# encoder (mbarrier pipelines, cross-set shared-memory hand-off, global atomicMax thresholds), its edge shapes and the
# tail-wave split, plus the sparse backward.  Run on a GPU box; logs land in gpurun_out/ (copy the summary to
# profiles/).   bash scripts/sanitize.sh [memcheck|racecheck|synccheck|initcheck ...]
set -uo pipefail
cd "$(dirname "${BASH_SOURCE[0]}")/.."
mkdir -p gpurun_out
TOOLS="${*:-memcheck racecheck synccheck}"
TESTS="tests/test_gpu_topk.py::test_fused_encoder_edge_shapes tests/test_gpu_topk.py::test_short_rows_and_ragged_sizes tests/test_gpu_topk.py::test_tail_wave_column_split_is_exact tests/test_gpu_topk.py::test_fused_fast_path_vs_oracle"
# memcheck also covers the other kernels that address shared / global memory by hand: the fused decode + dacts kernel
# (bulk copies into per-warp slots), the generic GEMM's TMA input / output boxes (L1 step, masked / store epilogues)
MEM_EXTRA="tests/test_gpu_topk.py::test_fused_decode_dacts_matches_separate_kernels tests/test_gpu_gemm.py tests/test_gpu_l1.py"
rc=0
for tool in $TOOLS; do
  log="gpurun_out/sanitize_${tool}.log"
  echo "== compute-sanitizer --tool $tool" | tee "$log"
  # -k filter keeps the racecheck pass (50-100x slowdown) to the small shapes
  filter=()
  [ "$tool" != "memcheck" ] && filter=(-k "not 19109 and not 6144")
  extra=""
  [ "$tool" = "memcheck" ] && extra="$MEM_EXTRA"
  timeout 1500 compute-sanitizer --tool "$tool" --target-processes all --error-exitcode 99 --print-limit 20 \
    python -m pytest $TESTS $extra -x -q -m gpu "${filter[@]}" >> "$log" 2>&1
  code=$?
  echo "exit code $code" | tee -a "$log"
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|error" "$log" | tail -5
  [ $code -ne 0 ] && rc=$code
done
exit $rc
