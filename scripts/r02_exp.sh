#!/bin/bash
mkdir -p gpurun_out
BUILDERS=host_sah,gpu_lbvh python scripts/gpu_builders.py 2>&1 | tail -4
echo "== greedy"; LMB200_GPU_GREEDY=1 BUILDERS=gpu_lbvh python scripts/gpu_builders.py 2>&1 | tail -2
for v in dpfull dpfull_ml1; do echo "== $v"; BUILDERS=gpu_lbvh,gpu_ploc LMB200_LIB=lightmetrica-v2_b200/lib/variants/liblmb200_$v.so python scripts/gpu_builders.py 2>&1 | tail -4; done
