#!/bin/bash
# ncu full capture of the traversal kernel (soup 4M, 16 Mi rays) and of k_extend (configs[2] scene)
mkdir -p gpurun_out
BENCH="python bench.py --rays 16777216 --steps 1 --warmup 3 --cpu-rays 100000 --pt-spp 8 --no-c4 --pt-cpu-spp 1"
ncu --set full --clock-control none --import-source on -k regex:trace_kernel -s 3 -c 1 -f -o gpurun_out/r02_prof_trace $BENCH --no-pt > gpurun_out/r02_prof_trace.log 2>&1
tail -3 gpurun_out/r02_prof_trace.log
ncu --set full --clock-control none --import-source on -k regex:k_extend -s 4 -c 1 -f -o gpurun_out/r02_prof_extend $BENCH > gpurun_out/r02_prof_extend.log 2>&1
tail -3 gpurun_out/r02_prof_extend.log
ls -la gpurun_out/*.ncu-rep
