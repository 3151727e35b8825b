#!/bin/bash
out=gpurun_out/${1:-r2h}; mkdir -p $out
export PIXIE_LIB_PATH=$PWD/ark_analysis_b200/_lib/libpixie_b200_prof.so
PIXIE_TC_STAGES=4 PIXIE_TRACE_STEP=20 timeout 120 python scripts/trace_train_step.py 5241600 32 10 10 > $out/trace_cfg2.log 2>&1
PIXIE_TC_STAGES=4 PIXIE_TRACE_STEP=4 timeout 120 python scripts/trace_train_step.py 5241600 32 10 10 > $out/trace_cfg2_early.log 2>&1
head -3 $out/trace_cfg2.log; head -3 $out/trace_cfg2_early.log
