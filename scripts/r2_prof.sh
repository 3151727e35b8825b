#!/bin/bash
out=gpurun_out/${1:-r2e}; mkdir -p $out
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,temperature.gpu --format=csv > $out/smi.log
for rep in 1 2; do
for shape in "5241600 32 10 10" "5241600 16 10 10" "5241600 64 10 10"; do
  echo "--- smem/default $shape" >> $out/prod.log
  timeout 120 python scripts/prof_train_pass.py $shape 5 >> $out/prod.log 2>&1
  echo "--- tabg=1 $shape" >> $out/prod.log
  PIXIE_TAB_GLOBAL=1 timeout 120 python scripts/prof_train_pass.py $shape 5 >> $out/prod.log 2>&1
done
done
echo "--- C=64 smem tables (NG=2)" >> $out/prod.log
PIXIE_TAB_GLOBAL=0 timeout 120 python scripts/prof_train_pass.py 5241600 64 10 10 5 >> $out/prod.log 2>&1
for shape in "3355392 40 20 20" "5000064 100 10 10" "26214400 40 20 20"; do
  timeout 120 python scripts/prof_train_pass.py $shape 5 >> $out/prod.log 2>&1
done
cat $out/prod.log
timeout 300 python bench.py --steps 10 --no-cpu > $out/bench.json 2> $out/bench.err; python -c "
import json;d=json.load(open('$out/bench.json'));print({k:d[k] for k in ['value','ms_per_step','train_ms_per_step','assign_ms_per_step','rows_rechecked_frac']}, d['roofline']['frac'])"
