#!/bin/bash
# round 2 (2 GPUs): the whole GPU suite with the multi-GPU tests actually running, then bench.py at N=1 and N=2
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r02_gpus.txt
python -m pytest tests -q -m gpu -rfs 2>&1 | tail -25 > gpurun_out/r02_pytest_gpu2.log
tail -8 gpurun_out/r02_pytest_gpu2.log
python bench.py --steps 3 --warmup 3 > gpurun_out/r02_bench1.json 2> gpurun_out/r02_bench1.err
tail -c 3000 gpurun_out/r02_bench1.json; tail -5 gpurun_out/r02_bench1.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r02_bench2.json 2> gpurun_out/r02_bench2.err
tail -c 2500 gpurun_out/r02_bench2.json; tail -5 gpurun_out/r02_bench2.err
