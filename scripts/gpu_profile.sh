#!/bin/bash
# ncu evidence for the traversal kernel: (1) launch list of a short bench run (per-launch device time,
# compare SHARES), (2) one full-set capture of the top kernel. Outputs under gpurun_out/.
set -x
mkdir -p gpurun_out
BENCH="python bench.py --rays 16777216 --steps 2 --warmup 3 --cpu-rays 100000 --pt-spp 32"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $BENCH > gpurun_out/bench_under_ncu.log 2>&1
# the default bench command's launch list (first 600 launches: all traversal steps + the start of the path-tracing part)
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_default.csv python bench.py --steps 2 --warmup 1 > gpurun_out/bench_default_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:trace_kernel -s 3 -c 1 -o gpurun_out/prof_trace $BENCH --no-pt > gpurun_out/prof_trace.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_extend -s 2 -c 1 -o gpurun_out/prof_extend $BENCH > gpurun_out/prof_extend.log 2>&1
ls -la gpurun_out/
