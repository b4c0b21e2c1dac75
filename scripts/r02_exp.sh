#!/bin/bash
mkdir -p gpurun_out
SWEEP_PT=0 python scripts/gpu_sweep.py > gpurun_out/r02_sweep2.txt 2>&1
cat gpurun_out/r02_sweep2.txt | cut -c1-400
