#!/bin/bash
# Baseline call of session 2: GPU tests, bench, mean-shift microbench, ncu of the persistent mean-shift kernel.
TAG=${1:-r1m}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_${TAG}.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_${TAG}.log
timeout 400 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
tail -c 600 gpurun_out/bench_${TAG}.json
timeout 200 python profiles/microbench_meanshift.py > gpurun_out/microbench_meanshift_${TAG}.txt 2>&1
cat gpurun_out/microbench_meanshift_${TAG}.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:mean_shift_fused_kernel -s 1 -c 1 -f -o gpurun_out/prof_${TAG}_msfused python profiles/microbench_meanshift.py > gpurun_out/ncu_${TAG}_msfused.log 2>&1
tail -2 gpurun_out/ncu_${TAG}_msfused.log
ls -la gpurun_out | grep ${TAG}
