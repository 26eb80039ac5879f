#!/bin/bash
# quick A/B session: sweep parity tests + a short bench of all workloads
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_batch.py tests/test_gpu_rce.py -m gpu -q -x 2>&1 | tail -2
timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('value',d['value'], 'kern_ms',d['roofline']['kernel_ms'], 'frac',d['roofline']['frac'], 'e2e',d['e2e']['ms_per_step'], d['e2e']['ms_per_step_with_rebuild'], 'rce',d['rce']['device_loop']['seconds'], d['rce']['host_loop']['seconds'])
for k,v in d.get('workloads',{}).items(): print(k, v.get('ms_per_step'), (v.get('roofline') or {}).get('frac'), (v.get('roofline') or {}).get('kernel_ms'))"
