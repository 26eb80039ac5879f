#!/bin/bash
# two GPUs: fused / separate exchange parity, then the default bench at N=2
set -u
mkdir -p gpurun_out
tag=${1:-r2n}
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q -s 2>&1 | grep -E "multi|passed|failed|Error" | tail -12
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu > gpurun_out/bench_n2_$tag.json 2> gpurun_out/bench_n2_$tag.err; echo "bench n2 rc=$?"; tail -4 gpurun_out/bench_n2_$tag.err
python - <<'PY'
import json,os
d=json.load(open('gpurun_out/bench_n2_%s.json' % os.environ.get('TAGV','r2n')))
print('N=2 value %.3e ms/step %.4f' % (d['value'], d['ms_per_step']))
for k,v in (d.get('workloads') or {}).items():
    print(k, {kk:vv for kk,vv in v.items() if kk in ('value','ms_per_step','scaling','n_gpus','max_rel_diff_vs_unsharded','sweep_kernel_ms')})
PY
