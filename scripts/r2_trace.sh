#!/bin/bash
out=gpurun_out/${1:-r2h}; mkdir -p $out
export PIXIE_LIB_PATH=$PWD/ark_analysis_b200/_lib/libpixie_b200_prof.so
PIXIE_TRACE_STEP=20 timeout 120 python scripts/trace_train_step.py 5241600 32 10 10 > $out/trace_cfg2.log 2>&1
PIXIE_TRACE_STEP=20 PIXIE_TAB_GLOBAL=1 timeout 120 python scripts/trace_train_step.py 5241600 32 10 10 > $out/trace_cfg2_tabg.log 2>&1
PIXIE_TRACE_STEP=20 timeout 120 python scripts/trace_train_step.py 3355392 40 20 20 > $out/trace_cfg3.log 2>&1
head -3 $out/trace_cfg2.log; wc -l $out/*.log
