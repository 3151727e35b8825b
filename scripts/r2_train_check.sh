#!/bin/bash
# round-2 training check: parity tests of the training paths, then timings of the three shapes
out=gpurun_out/${1:-r2b}; mkdir -p $out
timeout 900 python -m pytest tests/test_train_gpu.py tests/test_bmu_gpu.py -x -q > $out/pytest_train.log 2>&1; echo "rc=$?" >> $out/pytest_train.log
tail -15 $out/pytest_train.log
for shape in "5241600 32 10 10" "3355392 40 20 20" "5000064 100 10 10" "5241600 16 10 10" "5241600 64 10 10"; do
  timeout 120 python scripts/prof_train_pass.py $shape 5 >> $out/train_phases.log 2>&1
done
cat $out/train_phases.log
