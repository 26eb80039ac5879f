#!/bin/bash
# quick A/B session: sweep parity tests + a short C2 bench
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_batch.py -m gpu -q -x 2>&1 | tail -3
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --only-main 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('value',d['value'], 'kern_ms',d['roofline']['kernel_ms'], 'frac',d['roofline']['frac'], 'e2e',d['e2e']['ms_per_step'], d['e2e']['ms_per_step_with_rebuild'], 'rce',d['rce']['device_loop']['seconds'], d['rce']['host_loop']['seconds'])"
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-rce --workload C2 --batch 32 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('batch32 value',d['value'], 'kern_ms',d['roofline']['kernel_ms'], 'frac',d['roofline']['frac'])"
