#!/bin/bash
# ncu --set full of the ASSIGN kernel (plain variant) for cfg3's shape and for cfg2
out=gpurun_out/${1:-r2aa}; mkdir -p $out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:bmu_tc_kernel -s 1 -c 1 -o $out/assign_cfg3 python scripts/assign_stats.py 8 2048 40 20 20 > $out/ncu_cfg3.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:bmu_tc_kernel -s 1 -c 1 -o $out/assign_cfg2 python scripts/assign_stats.py 50 1024 32 10 10 > $out/ncu_cfg2.log 2>&1
tail -n 3 $out/ncu_cfg3.log; tail -n 3 $out/ncu_cfg2.log; ls -la $out
