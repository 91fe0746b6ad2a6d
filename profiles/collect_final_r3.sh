#!/bin/bash
# Final evidence of round 2, one GPU: full GPU test suite, smoke(), the default bench line, launch list of the bench command.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_final_r3.log 2>&1; tail -4 gpurun_out/pytest_final_r3.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
python bench.py > gpurun_out/bench_r3v.json 2> gpurun_out/bench_r3v.err; cut -c1-300 gpurun_out/bench_r3v.json
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_r3v.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-reference-config > gpurun_out/launches_r3v.log 2>&1
python profiles/summarize_launches.py gpurun_out/launches_r3v.csv 40 > gpurun_out/launches_r3v_summary.txt; head -14 gpurun_out/launches_r3v_summary.txt
