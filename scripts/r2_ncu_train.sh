#!/bin/bash
out=gpurun_out/${1:-r2g}; mkdir -p $out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:bmu_tc_kernel -s 2 -c 1 -o $out/train_cfg2 python scripts/prof_train_pass.py 5241600 32 10 10 1 > $out/ncu_train.log 2>&1
tail -5 $out/ncu_train.log
ls -la $out
