#!/bin/bash
# bench arms: ours (default line), reference (stock code path through oracle/refshim)
set -u
mkdir -p gpurun_out
tag=${1:-r2l}
timeout 900 python bench.py --impl reference > gpurun_out/bench_ref_$tag.json 2> gpurun_out/bench_ref_$tag.err; echo "bench ref rc=$?"; tail -3 gpurun_out/bench_ref_$tag.err
cat gpurun_out/bench_ref_$tag.json | cut -c1-1800
timeout 900 python bench.py > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo "bench rc=$?"; tail -5 gpurun_out/bench_$tag.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_%s.json' % "TAG".replace("TAG", __import__('os').environ.get('TAGV','r2l'))))
print('value %.3e ms/step %.4f  e2e ms %.4f (rebuild %.4f)  roofline frac %.3f kernel_ms %.4f launches %s' % (d['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['e2e']['ms_per_step_with_rebuild'], d['roofline']['frac'], d['roofline']['kernel_ms'], d['gpu_launches']))
print('rce', {k:(v if not isinstance(v,dict) else {kk:vv for kk,vv in v.items() if kk in ('seconds','radiation_iterations','convection_iterations')}) for k,v in (d.get('rce') or {}).items() if k!='what'})
print('cpu', d.get('cpu_baseline'))
for k,v in (d.get('workloads') or {}).items():
    if isinstance(v,dict):
        r=v.get('roofline') or {}
        print(k, 'ms', v.get('ms_per_step'), 'value', v.get('value'), 'frac', r.get('frac'), r.get('bound'), {kk:(vv.get('ms_per_species_loop') if isinstance(vv,dict) else vv) for kk,vv in v.items() if kk in ('RO','correlated-k')})
PY
timeout 300 python bench.py --impl reference --workload C3 > gpurun_out/bench_ref_c3_$tag.json 2> gpurun_out/bench_ref_c3_$tag.err; echo "bench ref C3 rc=$?"; cat gpurun_out/bench_ref_c3_$tag.json | cut -c1-600
