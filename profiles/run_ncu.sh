#!/bin/bash
# One `ncu --set full` capture per hot kernel (run under gpurun on ONE GPU; never wrap a multi-rank command).
# usage: bash profiles/run_ncu.sh <tag>      -> gpurun_out/prof_<tag>_<kernel>.ncu-rep
TAG=${1:-r1}
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline"
cap() {  # name regex skip count
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c $4 -f -o gpurun_out/prof_${TAG}_$1 $CMD > gpurun_out/ncu_${TAG}_$1.log 2>&1
  tail -1 gpurun_out/ncu_${TAG}_$1.log
}
cap mhsa mhsa_fwd_kernel 2 1
cap headmean attn_headmean_kernel 1 1
cap linear linear_tcgen05_kernel 3 3
cap ms_sim 'ms_sim' 2 1
cap ms_update 'ms_update' 1 1
