#!/bin/bash
# split-operand assign kernel: parity tests + same-box A/B against the plain kernel (PIXIE_X3=0)
out=gpurun_out/${1:-r2x3}; mkdir -p $out
timeout 900 python -m pytest tests/test_bmu_gpu.py tests/test_api_gpu.py -x -q > $out/pytest.log 2>&1; echo "rc=$?" >> $out/pytest.log
tail -3 $out/pytest.log
for rep in 1 2; do
 for shape in "50 1024 32 10 10" "50 1024 16 10 10" "50 1024 24 8 8" "50 1024 32 8 12"; do
  echo "--- plain: $shape" >> $out/ab.log
  PIXIE_X3=0 timeout 300 python scripts/assign_stats.py $shape 2>&1 | tail -1 >> $out/ab.log
  echo "--- split: $shape" >> $out/ab.log
  timeout 300 python scripts/assign_stats.py $shape 2>&1 | tail -1 >> $out/ab.log
 done
done
grep -v "^---" $out/ab.log | awk '{print $1,$2,$5,$6,$10,$11,$12,$13,$14,$15}' | paste - - 
