#!/bin/bash
# one --set full capture of each shading kernel of the wavefront (steady state: skip the first iterations)
mkdir -p gpurun_out
BENCH="python bench.py --rays 1048576 --steps 1 --warmup 3 --cpu-rays 10000 --pt-spp 32 --no-c4 --no-one --pt-cpu-spp 1"
ncu --set full --clock-control none --import-source on -k regex:'k_logic|k_bsdf|k_nee|k_shadow' -s 40 -c 4 -f -o gpurun_out/r02_prof_shading $BENCH > gpurun_out/r02_prof_shading.log 2>&1
ls -la gpurun_out/r02_prof_shading*
