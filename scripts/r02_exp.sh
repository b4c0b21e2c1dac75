#!/bin/bash
for c in 23 24 25; do echo "chunk 2^$c"; LMB200_E2E_CHUNK_LOG2=$c python bench.py --steps 3 --warmup 3 --no-pt --no-c4 --no-one 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('value',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'full',round(d['e2e']['full_ray_form']['value'],1),'frac_dev',round(d['e2e']['frac_of_device_rate'],3))"; done
