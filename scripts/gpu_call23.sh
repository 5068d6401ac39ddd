#!/bin/bash
set -u
O=gpurun_out/c23
mkdir -p $O
CUDA_LAUNCH_BLOCKING=1 timeout 300 python scripts/aux_prof.py c3 > $O/aux_block.log 2>&1; grep -n "failed\|RuntimeError" $O/aux_block.log | head -5; tail -3 $O/aux_block.log | cut -c1-200
FREUD_NO_TMA_OUT=1 timeout 300 python scripts/aux_prof.py c3 > $O/aux_notma.log 2>&1; head -12 $O/aux_notma.log | cut -c1-200; tail -1 $O/aux_notma.log
timeout 300 python scripts/aux_prof.py c3 > $O/aux.log 2>&1; head -12 $O/aux.log | cut -c1-200; tail -1 $O/aux.log
timeout 300 python scripts/aux_prof.py c2 > $O/aux_c2.log 2>&1; head -8 $O/aux_c2.log | cut -c1-200; tail -1 $O/aux_c2.log
