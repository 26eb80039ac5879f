#!/bin/bash
# the whole GPU suite (with durations) and the default bench
set -u
mkdir -p gpurun_out
tag=${1:-full}
timeout 1800 python -m pytest tests -m gpu -q --durations=12 > gpurun_out/pytest_$tag.log 2>&1; echo "pytest full rc=$?"; tail -30 gpurun_out/pytest_$tag.log
timeout 900 python bench.py > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_$tag.err
python - <<'PY'
import json,os
d=json.load(open('gpurun_out/bench_%s.json' % os.environ.get('TAGV','full')))
print('value %.3e ms/step %.4f  e2e ms %.4f (rebuild %.4f)  roofline frac %.3f kernel_ms %.4f launches %s' % (d['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['e2e']['ms_per_step_with_rebuild'], d['roofline']['frac'], d['roofline']['kernel_ms'], d['gpu_launches']))
for k,v in (d.get('workloads') or {}).items():
    if isinstance(v,dict):
        r=v.get('roofline') or {}
        print(k, 'ms', v.get('ms_per_step'), 'kernel', v.get('sweep_kernel_ms'), 'value', v.get('value'), 'frac', r.get('frac'), r.get('bound'))
PY
