#!/bin/bash
# parity (kernels, batch bit-identity, full size) + the default bench without the CPU / convergence legs
set -u
mkdir -p gpurun_out
tag=${1:-q}
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_batch.py -m gpu -q -x > gpurun_out/pytest_$tag.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_$tag.log
timeout 900 python bench.py --no-cpu --only-main > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_$tag.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_$tag.json'))
print('value %.3e ms/step %.4f  e2e ms %.4f (rebuild %.4f)  roofline frac %.3f kernel_ms %.4f' % (d['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['e2e']['ms_per_step_with_rebuild'], d['roofline']['frac'], d['roofline']['kernel_ms']))
r=d.get('rce') or {}
print({k:(v.get('seconds'), v.get('wall_breakdown')) for k,v in r.items() if isinstance(v,dict) and 'seconds' in v})
PY
