#!/bin/bash
python -m pytest tests/test_gpu_trace.py -q -x -k "soup_vs_oracle or mesh_scene or edge or axis" 2>&1 | tail -2
LMB200_LIB=lightmetrica-v2_b200/lib/variants/liblmb200_pipe.so python -m pytest tests/test_gpu_trace.py tests/test_gpu_render.py -q -x -k "soup_vs_oracle or mesh_scene or edge or axis or same_samples" 2>&1 | tail -2
python scripts/gpu_sweep.py 2>&1 | cut -c1-330
