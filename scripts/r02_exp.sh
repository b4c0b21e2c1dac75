#!/bin/bash
SWEEP_BUILDER=0 python scripts/gpu_sweep.py hgreedy 2>&1 | cut -c1-420
