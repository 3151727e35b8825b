#!/bin/bash
# runs each stage of scripts/gpu_check.py in its own process with a timeout (a trapped kernel
# poisons its CUDA context, so stages are isolated).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv | tee gpurun_out/check.log
for st in "$@"; do
  timeout 300 python scripts/gpu_check.py $st 2>&1 | tail -40 | tee -a gpurun_out/check.log
  echo "stage $st exit ${PIPESTATUS[0]}" | tee -a gpurun_out/check.log
done
