#!/bin/bash
# quick visit: GPU parity suite + headline numbers
mkdir -p gpurun_out
python -m pytest tests -q -m gpu -x -rfs 2>&1 | tail -15 > gpurun_out/r02_pytest_quick.log
tail -4 gpurun_out/r02_pytest_quick.log
python bench.py --steps 3 --warmup 3 --pt-spp 64 > gpurun_out/r02_quick.json 2> gpurun_out/r02_quick.err
tail -3 gpurun_out/r02_quick.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_quick.json').read().strip().splitlines()[-1])
r=d['roofline']; pt=d['path_tracing']
print('soup Mrays/s', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'nodes/ray', round(r['nodes_per_ray'],2), 'tris/ray', round(r['tris_per_ray'],2), 'frac', round(r['frac'],3), 'parity', d['parity'], d['bvh'])
print('intersect_one', d['intersect_one'])
print('mesh Mrays/s', round(pt['incoherent_1m_tri_mesh']['value'],1), 'pt Msamples/s', round(pt['value'],1), 'pt e2e', round(pt['e2e']['value'],1), 'pt roof', pt['roofline']['frac'])
print('pt cpu', pt['cpu_baseline'])
print('c4', d['config4']['value'], d['config4']['bvh'])
PY
