#!/bin/bash
# split-operand assign kernel: parity tests, window-margin sweep, same-box A/B against the plain kernel
out=gpurun_out/${1:-r2x3}; mkdir -p $out
timeout 900 python -m pytest tests/test_bmu_gpu.py tests/test_api_gpu.py -x -q > $out/pytest.log 2>&1; echo "rc=$?" >> $out/pytest.log
tail -5 $out/pytest.log
timeout 600 python scripts/x3_margin.py 8 > $out/margin.log 2>&1; tail -45 $out/margin.log
for rep in 1 2; do
 for shape in "50 1024 32 10 10" "50 1024 16 10 10" "50 1024 24 8 8"; do
  echo "--- plain: $shape" >> $out/ab.log
  PIXIE_X3=0 timeout 300 python scripts/assign_stats.py $shape 2>&1 | tail -1 >> $out/ab.log
  echo "--- split: $shape" >> $out/ab.log
  timeout 300 python scripts/assign_stats.py $shape 2>&1 | tail -1 >> $out/ab.log
 done
done
cat $out/ab.log
