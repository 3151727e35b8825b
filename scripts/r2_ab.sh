#!/bin/bash
out=gpurun_out/${1:-r2s}; mkdir -p $out
R1=$PWD/ark_analysis_b200/_lib/libpixie_b200_r1.so
for rep in 1 2; do
  shape="5241600 32 10 10"
  echo "--- r1 $shape" >> $out/ab.log
  PIXIE_LIB_PATH=$R1 timeout 120 python scripts/prof_train_pass.py $shape 5 >> $out/ab.log 2>&1
  echo "--- now (8 stages) $shape" >> $out/ab.log
  timeout 120 python scripts/prof_train_pass.py $shape 5 >> $out/ab.log 2>&1
  echo "--- now, 4 stages $shape" >> $out/ab.log
  PIXIE_TC_STAGES=4 timeout 120 python scripts/prof_train_pass.py $shape 5 >> $out/ab.log 2>&1
  echo "--- now, tabg 8 stages $shape" >> $out/ab.log
  PIXIE_TAB_GLOBAL=1 timeout 120 python scripts/prof_train_pass.py $shape 5 >> $out/ab.log 2>&1
  echo "--- now, tabg 4 stages $shape" >> $out/ab.log
  PIXIE_TC_STAGES=4 PIXIE_TAB_GLOBAL=1 timeout 120 python scripts/prof_train_pass.py $shape 5 >> $out/ab.log 2>&1
done
cat $out/ab.log
