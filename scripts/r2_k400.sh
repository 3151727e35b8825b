#!/bin/bash
out=gpurun_out/${1:-r2ac}; mkdir -p $out
timeout 900 python -m pytest tests/test_bmu_gpu.py tests/test_train_gpu.py -x -q > $out/pytest.log 2>&1; echo "rc=$?" >> $out/pytest.log
tail -n 5 $out/pytest.log
for a in "8 1024 32 20 20" "8 1024 16 20 20" "8 1024 24 18 18"; do
  echo "--- default $a" >> $out/k400.log
  timeout 200 python scripts/assign_stats.py $a >> $out/k400.log 2>&1
  echo "--- forced 100,2,2,2 / 104,2,2,2 $a" >> $out/k400.log
  PIXIE_TC_VARIANT=100,2,2,2 timeout 200 python scripts/assign_stats.py $a >> $out/k400.log 2>&1
  PIXIE_TC_VARIANT=104,2,2,2 timeout 200 python scripts/assign_stats.py $a >> $out/k400.log 2>&1
done
grep -v Warn $out/k400.log
