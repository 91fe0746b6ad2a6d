#!/bin/bash
# Round-1 final artefacts in one GPU call:  bash profiles/collect_final.sh <tag>
TAG=${1:-r1z}
mkdir -p gpurun_out
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
cut -c1-300 gpurun_out/bench_${TAG}.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_${TAG}_list.log 2>&1
cap() {  # name kernel-regex skip script args...
  local name=$1 rx=$2 skip=$3; shift 3
  timeout 200 ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c 1 -f -o gpurun_out/prof_${TAG}_$name "$@" > gpurun_out/ncu_${TAG}_$name.log 2>&1
  tail -1 gpurun_out/ncu_${TAG}_$name.log
}
cap mhsa mhsa_fwd2_kernel 1 python profiles/ncu_targets.py mhsa
cap headmean attn_headmean2_kernel 1 python profiles/ncu_targets.py headmean
cap fc1 linear_tcgen05_kernel 1 python profiles/ncu_targets.py fc1
cap proj linear_tcgen05_kernel 1 python profiles/ncu_targets.py proj
cap msfused mean_shift_fused_kernel 1 python profiles/microbench_meanshift.py
python profiles/microbench.py > gpurun_out/microbench_${TAG}.txt 2>&1
python profiles/microbench_meanshift.py > gpurun_out/microbench_meanshift_${TAG}.txt 2>&1
python profiles/microbench_mhsa.py > gpurun_out/microbench_mhsa_${TAG}.txt 2>&1
ls gpurun_out | grep ${TAG}
