#!/bin/bash
# Bench lines + launch list of the current build (no ncu --set full captures): refreshes the numbers under profiles/
# after a change that does not touch the profiled kernels' code.   bash scripts/capture_bench_only.sh r02
set -u
TAG=${1:-r02}
OUT=gpurun_out/cap_$TAG
mkdir -p $OUT
python bench.py --profile-out $OUT/prof_c3.json > $OUT/bench_c3.json 2> $OUT/bench_c3.err
python bench.py --workload c2 --no-cpu-baseline --no-eager --profile-out $OUT/prof_c2.json > $OUT/bench_c2.json 2> $OUT/bench_c2.err
python bench.py --workload c2 --precision fp32 --no-cpu-baseline --no-eager --no-parity > $OUT/bench_c2_fp32.json 2> /dev/null
python bench.py --workload c4 --no-cpu-baseline --no-eager --no-parity --profile-out $OUT/prof_c4.json > $OUT/bench_c4_1gpu.json 2> $OUT/bench_c4.err
python scripts/bench_extra.py > $OUT/bench_extra.json 2> $OUT/bench_extra.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_c3.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-eager --no-parity --no-extras > $OUT/ncu_launch.log 2>&1
cuobjdump -sass freud_b200/libfreud_b200.so | grep -oE "UTCHMMA[A-Z0-9_.]*|UTCQMMA[A-Z0-9_.]*|LDTM[A-Z0-9_.]*|STTM[A-Z0-9_.]*|UTMALDG[A-Z0-9_.]*|UTMASTG[A-Z0-9_.]*|UBLKCP[A-Z0-9_.]*|UBLKPF[A-Z0-9_.]*|UTCBAR[A-Z0-9_.]*|SYNCS[A-Z0-9_.]*|LDGMC[A-Z0-9_.]*|HMMA[A-Z0-9_.]*|REDUX[A-Z0-9_.]*|FFMA2" | sort | uniq -c | sort -rn > $OUT/sass_mnemonics.txt
FREUD_ENC_STATS=1 python scripts/enc_stats.py > $OUT/enc_stats.txt 2>&1
ls -la $OUT
