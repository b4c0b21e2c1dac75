#!/bin/bash
# round-2 ncu evidence: launch lists of the bench commands (compare SHARES) + one full-set capture of trace_kernel and k_extend
mkdir -p gpurun_out
BENCH="python bench.py --rays 16777216 --steps 2 --warmup 3 --cpu-rays 100000 --pt-spp 32 --no-c4 --no-one --pt-cpu-spp 1"
ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/r02_launches.csv $BENCH > gpurun_out/r02_bench_under_ncu.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_default.csv python bench.py --steps 2 --warmup 3 --no-c4 --no-one > gpurun_out/r02_bench_default_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:trace_kernel -s 3 -c 1 -f -o gpurun_out/r02_prof_trace $BENCH --no-pt > gpurun_out/r02_prof_trace.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_extend -s 4 -c 1 -f -o gpurun_out/r02_prof_extend $BENCH > gpurun_out/r02_prof_extend.log 2>&1
ls -la gpurun_out/r02_*
