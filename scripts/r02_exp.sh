#!/bin/bash
python -m pytest tests -q -m gpu -x 2>&1 | tail -3
python scripts/gpu_sweep.py 2>&1 | cut -c1-330
