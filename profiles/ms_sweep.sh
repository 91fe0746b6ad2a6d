#!/bin/bash
# stagger sweep + phase tables of the mean-shift kernels (run on the GPU box)
cd "$(dirname "$0")/.."
python profiles/occ_probe.py
for st in ${STAGGERS:-0 12000}; do
  echo "=== AS_MS_STAGGER_NS=$st"
  AS_MS_STAGGER_NS=$st python profiles/microbench_meanshift.py 2>&1 | grep -E "^v2 |^fused |v2 total|^[0-9]+ .*(us|  )" | head -16
done
