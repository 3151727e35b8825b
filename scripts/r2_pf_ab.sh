#!/bin/bash
# same-box A/B of the L2 prefetch (PIXIE_DBG_FLAGS=8 switches it off) on the assign and training
# launches, after the parity tests of both paths
out=gpurun_out/${1:-r2pf}; mkdir -p $out
timeout 900 python -m pytest tests/test_train_gpu.py tests/test_bmu_gpu.py -x -q > $out/pytest.log 2>&1; echo "rc=$?" >> $out/pytest.log
tail -4 $out/pytest.log
for rep in 1 2; do
 for shape in "50 1024 32 10 10" "8 2048 40 20 20" "8 1024 16 20 20" "5 1024 100 10 10"; do
  echo "--- prefetch off: $shape" >> $out/ab.log
  PIXIE_DBG_FLAGS=8 timeout 300 python scripts/assign_stats.py $shape 2>&1 | tail -1 >> $out/ab.log
  echo "--- prefetch on: $shape" >> $out/ab.log
  timeout 300 python scripts/assign_stats.py $shape 2>&1 | tail -1 >> $out/ab.log
 done
 for shape in "5241600 32 10 10" "3355392 40 20 20" "5000064 100 10 10"; do
  echo "--- train, prefetch off: $shape" >> $out/ab.log
  PIXIE_DBG_FLAGS=8 timeout 120 python scripts/prof_train_pass.py $shape 5 2>&1 | tail -2 >> $out/ab.log
  echo "--- train, prefetch on: $shape" >> $out/ab.log
  timeout 120 python scripts/prof_train_pass.py $shape 5 2>&1 | tail -2 >> $out/ab.log
 done
done
cat $out/ab.log
