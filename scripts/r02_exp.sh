#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_trace.py -q -x -k "service or single_ray" 2>&1 | tail -3
python -m pytest tests/test_gpu_plugin.py -q -x 2>&1 | tail -3
python scripts/r02_service.py 2>&1 | tail -4 | tee gpurun_out/r02_service.txt
