#!/bin/bash
# bench.py at N=2 and N=4 on a 4-GPU lease (the driver's SCALE run does 1, 2, 4, 8 itself; this fills profiles/r02_multi_gpu.md)
mkdir -p gpurun_out
for n in 2 4; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2952$n bench.py --gpus $n --steps 5 --warmup 3 > gpurun_out/r02_bench$n.json 2> gpurun_out/r02_bench$n.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r02_bench$n.json').read().strip().splitlines()[-1])
pt=d['path_tracing']; c4=d['config4']
print('N=$n value',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'pcie',round(d['e2e']['pcie_ceiling_gbs'],1),'frac',round(d['e2e']['frac_of_pcie_ceiling'],3),'full',round(d['e2e']['full_ray_form']['value'],1))
print('   pt',round(pt['value'],1),'e2e',round(pt['e2e']['value'],1),'mesh',round(pt['incoherent_1m_tri_mesh']['value'],1),'c4',round(c4['value'],1),'reduce_ms',c4['film_reduce_ms'],'render_ms',round(c4['render_ms'],1))
PY
done
