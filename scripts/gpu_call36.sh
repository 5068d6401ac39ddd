#!/bin/bash
set -u
O=gpurun_out/c36
mkdir -p $O
ONLY_SHAPE=75000,384,6144 ncu --set full --clock-control none --import-source on -k regex:sm100_topk_kernel -s 3 -c 1 -o $O/enc_c2 python scripts/enc_variants.py > $O/ncu.log 2>&1
ncu -i $O/enc_c2.ncu-rep --page source --csv > $O/enc_c2.source.csv 2> /dev/null
ncu -i $O/enc_c2.ncu-rep --page raw --csv > $O/enc_c2.raw.csv 2> /dev/null
rm -f $O/enc_c2.ncu-rep
ls -la $O
