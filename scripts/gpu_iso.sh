#!/bin/bash
# quick look at the isothermal flux-sweep kernel: C1 single, C5 batch, C4 spectrum
for w in C1 C5 C4; do
  timeout 600 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu --no-rce --only-main 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('$w value %.3e ms_per_step %.4f kernel_ms %.4f frac %.3f' % (d['value'], d['ms_per_step'], r['kernel_ms'], r['frac']))"
done
