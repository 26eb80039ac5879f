#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_fband_lane -s 4 -c 1 \
    -o gpurun_out/prof_lane -f python bench.py --steps 3 --warmup 3 --no-cpu --no-rce --only-main > gpurun_out/ncu_lane.log 2>&1
echo "ncu rc=$?"
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --only-main 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('value',d['value'], 'kern_ms',d['roofline']['kernel_ms'], 'rce',d['rce'])"
