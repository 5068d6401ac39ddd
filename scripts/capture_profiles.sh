#!/bin/bash
# Round artefacts, captured on the GPU box (gpurun -- 'bash scripts/capture_profiles.sh r01').
# Bench lines are taken WITHOUT a profiler; ncu runs are separate processes.  Everything lands in gpurun_out/cap_<tag>/;
# scripts/summarize_profiles.py turns that directory into the tracked files under profiles/.
set -u
TAG=${1:-r01}
OUT=gpurun_out/cap_$TAG
mkdir -p $OUT
python bench.py --profile-out $OUT/prof_c3.json > $OUT/bench_c3.json 2> $OUT/bench_c3.err
python bench.py --workload c2 --no-cpu-baseline --profile-out $OUT/prof_c2.json > $OUT/bench_c2.json 2> $OUT/bench_c2.err
python bench.py --workload c2 --precision fp32 --no-cpu-baseline > $OUT/bench_c2_fp32.json 2> /dev/null
python bench.py --workload c4 --no-cpu-baseline --profile-out $OUT/prof_c4.json > $OUT/bench_c4_1gpu.json 2> $OUT/bench_c4.err
python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_reference_arm.json 2> $OUT/bench_ref.err
python scripts/bench_extra.py > $OUT/bench_extra.json 2> $OUT/bench_extra.err
# launch list of the bench step (cold-cache, serialised: shares only)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_c3.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/ncu_launch.log 2>&1
# one --set full capture each of the dominant kernels (launch of the 4th step)
for spec in "enc:sm100_gemm_kernel" "sgw:sparse_grads_warp_kernel" "dec:decode_fixed_kernel" "dac:dacts_fixed_kernel" "adam:adam_kernel"; do
  name=${spec%%:*}; rx=${spec##*:}
  ncu --set full --clock-control none --import-source on -k regex:$rx -s 3 -c 1 -o $OUT/full_${name}_c3 \
      python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $OUT/ncu_full_$name.log 2>&1
done
ls -la $OUT
