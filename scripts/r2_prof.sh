#!/bin/bash
out=gpurun_out/${1:-r2e}; mkdir -p $out
timeout 600 python -m pytest tests/test_train_gpu.py -x -q > $out/pytest_train.log 2>&1; echo "rc=$?" >> $out/pytest_train.log
tail -3 $out/pytest_train.log
for shape in "5241600 32 10 10" "3355392 40 20 20" "5000064 100 10 10" "26214400 40 20 20" "5241600 64 10 10" "5241600 16 10 10"; do
  timeout 120 python scripts/prof_train_pass.py $shape 5 >> $out/prod.log 2>&1
done
echo "--- C=64 smem tables (NG=2)" >> $out/prod.log
PIXIE_TAB_GLOBAL=0 timeout 120 python scripts/prof_train_pass.py 5241600 64 10 10 5 >> $out/prod.log 2>&1
cat $out/prod.log
export PIXIE_LIB_PATH=$PWD/ark_analysis_b200/_lib/libpixie_b200_prof.so
PIXIE_TRACE_STEP=20 timeout 120 python scripts/trace_train_step.py 5241600 32 10 10 > $out/trace_cfg2.log 2>&1
PIXIE_TRACE_STEP=20 timeout 120 python scripts/trace_train_step.py 3355392 40 20 20 > $out/trace_cfg3.log 2>&1
