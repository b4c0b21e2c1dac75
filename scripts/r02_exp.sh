#!/bin/bash
python -m pytest tests/test_gpu_trace.py -q -x -k "deep_device" 2>&1 | tail -30
