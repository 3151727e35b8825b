#!/bin/bash
# tail8 layout (32-byte rows for the last 8 channels): parity tests, then same-box A/B (PIXIE_TAIL8=0 = off)
out=gpurun_out/${1:-r2t8}; mkdir -p $out
timeout 1200 python -m pytest tests/test_bmu_gpu.py tests/test_train_gpu.py tests/test_api_gpu.py -x -q > $out/pytest.log 2>&1; echo "rc=$?" >> $out/pytest.log
tail -8 $out/pytest.log
for rep in 1 2; do
 for shape in "8 2048 40 20 20" "5 1024 100 10 10" "20 1024 40 10 10" "20 1024 72 10 10"; do
  echo "off: $(PIXIE_TAIL8=0 timeout 300 python scripts/assign_stats.py $shape 2>&1 | tail -1)" >> $out/ab.log
  echo "on : $(timeout 300 python scripts/assign_stats.py $shape 2>&1 | tail -1)" >> $out/ab.log
 done
 for shape in "3355392 40 20 20" "5000064 100 10 10"; do
  echo "off: $(PIXIE_TAIL8=0 python scripts/prof_train_pass.py $shape 5 2>&1 | tail -2 | head -1)" >> $out/ab.log
  echo "on : $(python scripts/prof_train_pass.py $shape 5 2>&1 | tail -2 | head -1)" >> $out/ab.log
 done
done
cut -c1-150 $out/ab.log
