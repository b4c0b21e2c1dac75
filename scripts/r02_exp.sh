#!/bin/bash
python -m pytest tests/test_gpu_plugin.py -q -rf 2>&1 | tail -5
