#!/bin/bash
# One GPU visit: parity tests, smoke, bench, ncu launch list (shares) — outputs under gpurun_out/.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt
nproc > gpurun_out/host.txt; lscpu | grep "Model name" >> gpurun_out/host.txt
python -m pytest tests -q -m gpu 2>&1 | tail -30 > gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -3 gpurun_out/pytest_gpu.log; cat gpurun_out/smoke.log | tail -2; cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
