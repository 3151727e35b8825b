#!/bin/bash
# one process per configuration (a trapped kernel poisons its CUDA context)
mkdir -p gpurun_out
L=gpurun_out/variants.log
: > $L
run() { echo "--- $*" | tee -a $L; ( export "${@:1:$#-1}"; timeout 300 python scripts/variant_experiment.py ${!#} ) 2>&1 | grep -v Warning | tail -8 | tee -a $L; }
run X=1 "parity 32 100 2"
run X=1 "timing 32 100 50"
run PIXIE_DELTA_SCALE=0.0001 "timing 32 100 50"
run X=1 "timingU 32 100 50"
run X=1 "timing 16 100 50"
run X=1 "timing 40 400 24"
run X=1 "timing 100 100 5"
run X=1 "timing 64 100 24"
