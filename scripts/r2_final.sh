#!/bin/bash
# end-of-round pass: full GPU test-suite, the bench line at N=1 (and N=2 when two GPUs are visible),
# the reference arm.  Output under gpurun_out/r2final_*.
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l)
python -m pytest tests -q -m gpu -x > gpurun_out/r2final_pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/r2final_pytest_gpu.log
tail -n 3 gpurun_out/r2final_pytest_gpu.log
python bench.py > gpurun_out/r2final_bench_n1.json 2> gpurun_out/r2final_bench_n1.err; tail -c 1500 gpurun_out/r2final_bench_n1.json
if [ "$NG" -ge 2 ]; then
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 2 > gpurun_out/r2final_bench_n2.json 2> gpurun_out/r2final_bench_n2.err
  tail -c 600 gpurun_out/r2final_bench_n2.json
fi
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2final_bench_ref.json 2> gpurun_out/r2final_bench_ref.err
cat gpurun_out/r2final_bench_ref.json
