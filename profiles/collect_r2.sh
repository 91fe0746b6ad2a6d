#!/bin/bash
# Round-2 evidence, one GPU: bench line, launch list of the same command, ncu --set full of the kernels VERDICT named.
#   bash profiles/collect_r2.sh <tag> [bench|launches|ncu ...]
cd "$(dirname "$0")/.."
TAG=$1; shift
mkdir -p gpurun_out
for W in "$@"; do
  case $W in
    bench)
      python bench.py --steps 10 --warmup 3 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; tail -c 3000 gpurun_out/bench_${TAG}.json;;
    launches)
      timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_${TAG}.csv \
        python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-reference-config > gpurun_out/launches_${TAG}.log 2>&1
      python profiles/summarize_launches.py gpurun_out/launches_${TAG}.csv 40 > gpurun_out/launches_${TAG}_summary.txt; head -30 gpurun_out/launches_${TAG}_summary.txt;;
    ncu_ms)
      timeout 300 ncu --set full --clock-control none --import-source on -k regex:mean_shift_v2 -s 1 -c 1 -f -o gpurun_out/prof_${TAG}_msv2 python profiles/ncu_ms_target.py v2 > gpurun_out/ncu_${TAG}_msv2.log 2>&1; tail -1 gpurun_out/ncu_${TAG}_msv2.log;;
    ncu_msf)
      timeout 300 ncu --set full --clock-control none --import-source on -k regex:mean_shift_fused -s 1 -c 1 -f -o gpurun_out/prof_${TAG}_msfused python profiles/ncu_ms_target.py fused > gpurun_out/ncu_${TAG}_msfused.log 2>&1; tail -1 gpurun_out/ncu_${TAG}_msfused.log;;
    ncu_hm)
      for V in headmean headmean_lean headmean_rows; do
        timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_headmean2 -s 1 -c 1 -f -o gpurun_out/prof_${TAG}_$V python profiles/ncu_targets.py $V > gpurun_out/ncu_${TAG}_$V.log 2>&1; tail -1 gpurun_out/ncu_${TAG}_$V.log
      done;;
    ncu_bwd)
      for K in mhsa_bwd_dkv mhsa_bwd_dq; do
        timeout 300 ncu --set full --clock-control none --import-source on -k regex:$K -s 1 -c 1 -f -o gpurun_out/prof_${TAG}_$K python profiles/ncu_targets.py mhsa_bwd > gpurun_out/ncu_${TAG}_$K.log 2>&1; tail -1 gpurun_out/ncu_${TAG}_$K.log
      done;;
    ncu_mhsa)
      timeout 300 ncu --set full --clock-control none --import-source on -k regex:mhsa_fwd2 -s 1 -c 1 -f -o gpurun_out/prof_${TAG}_mhsa python profiles/ncu_targets.py mhsa > gpurun_out/ncu_${TAG}_mhsa.log 2>&1; tail -1 gpurun_out/ncu_${TAG}_mhsa.log;;
  esac
done
