#!/bin/bash
# Everything the round's numbers come from, in one GPU call:  bash profiles/collect.sh <tag>
TAG=${1:-r1}
mkdir -p gpurun_out
timeout 500 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
tail -c 400 gpurun_out/bench_${TAG}.json
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_${TAG}_list.log 2>&1
bash profiles/run_ncu2.sh ${TAG} mhsa headmean qkv
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline"
for K in tc_update tc_zpart ccl_init_runs; do
  timeout 200 ncu --set full --clock-control none --import-source on -k regex:$K -s 1 -c 1 -f -o gpurun_out/prof_${TAG}_$K $CMD > gpurun_out/ncu_${TAG}_$K.log 2>&1
done
python profiles/microbench.py > gpurun_out/microbench_${TAG}.txt 2>&1
ls gpurun_out | grep ${TAG}
