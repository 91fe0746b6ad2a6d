#!/bin/bash
# ncu --set full of individual kernels at cfg2 shapes.  usage: bash profiles/run_ncu2.sh <tag> <which...>
TAG=$1; shift
for W in "$@"; do
  case $W in
    proj|fc1|qkv) K=linear_tcgen05_kernel; S=1;;
    mhsa) K=mhsa_fwd_kernel; S=1;;
    headmean) K=attn_headmean_kernel; S=1;;
  esac
  [ $W = qkv ] && S=1
  [ $W = mhsa ] && S=1
  timeout 200 ncu --set full --clock-control none --import-source on -k regex:$K -s $S -c 1 -f -o gpurun_out/prof_${TAG}_$W python profiles/ncu_targets.py $W > gpurun_out/ncu_${TAG}_$W.log 2>&1
  tail -1 gpurun_out/ncu_${TAG}_$W.log
done
