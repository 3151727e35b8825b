#!/bin/bash
# same-box A/B of several builds of the library on the assign launch: args after the output name
# are library files under ark_analysis_b200/_lib (selected through PIXIE_LIB_PATH); "-" = the
# tree's own library.  Parity tests of the assign path (tree's library) first.
out=gpurun_out/${1:-r2dyn}; mkdir -p $out; shift
LIBS=${@:-"libpixie_b200_pf.so -"}
timeout 900 python -m pytest tests/test_bmu_gpu.py tests/test_api_gpu.py -x -q > $out/pytest.log 2>&1; echo "rc=$?" >> $out/pytest.log
tail -4 $out/pytest.log
for rep in 1 2; do
 for shape in "50 1024 32 10 10" "50 1024 16 10 10" "8 2048 40 20 20" "5 1024 100 10 10" "20 1024 64 10 10"; do
  for lib in $LIBS; do
   echo "--- $lib: $shape" >> $out/ab.log
   if [ "$lib" = "-" ]; then
     timeout 300 python scripts/assign_stats.py $shape 2>&1 | tail -1 >> $out/ab.log
   else
     PIXIE_LIB_PATH=$PWD/ark_analysis_b200/_lib/$lib timeout 300 python scripts/assign_stats.py $shape 2>&1 | tail -1 >> $out/ab.log
   fi
  done
 done
done
grep -v "^---" $out/ab.log | awk '{print $1,$2,$5,$6,$10,$11,$12,$13}' | paste - - - | head -40
