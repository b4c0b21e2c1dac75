#!/bin/bash
# e2e of the host-buffer call: streaming launch (default) against the per-chunk launches (LMB200_E2E_STREAM=0)
for cfg in "1 0" "0 0" "1 0" "1 256"; do
  set -- $cfg
  LMB200_E2E_STREAM=$1 LMB200_E2E_GRADE=$2 python bench.py --steps 6 --warmup 3 --no-pt --no-c4 --no-one --cpu-rays 100000 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('stream $1 grade $2 value %.1f e2e %.1f (%.3f of device) full %.1f parity %s' % (d['value'], d['e2e']['value'], d['e2e']['frac_of_device_rate'], d['e2e']['full_ray_form']['value'], d['parity']['index_mismatches']))"
done
LMB200_STREAM_DEBUG=1 python bench.py --steps 1 --warmup 3 --no-pt --no-c4 --no-one --cpu-rays 10000 2>&1 >/dev/null | grep "stream\]" | tail -30 | cut -c1-100
