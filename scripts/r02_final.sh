#!/bin/bash
# the driver's round-end sequence on one GPU: GPU suite, smoke, reference arm, product arm (default flags)
mkdir -p gpurun_out
python -m pytest tests -q -m gpu -rfs 2>&1 | tail -10 > gpurun_out/r02_pytest_gpu1.log
tail -3 gpurun_out/r02_pytest_gpu1.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1; tail -2 gpurun_out/r02_smoke.log
( time python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02_bench_ref.json 2> gpurun_out/r02_bench_ref.err ) 2>&1 | grep real
( time python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02_bench1.json 2> gpurun_out/r02_bench1.err ) 2>&1 | grep real
tail -2 gpurun_out/r02_bench1.err
cut -c1-1500 gpurun_out/r02_bench_ref.json
