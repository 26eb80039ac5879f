#!/bin/bash
# the default bench at N GPUs (C2 replicas + the partitioned workloads C4 / C5)
set -u
mkdir -p gpurun_out
N=${1:-2}
tag=${2:-r2}
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 20 --warmup 3 --no-cpu > gpurun_out/bench_n${N}_$tag.json 2> gpurun_out/bench_n${N}_$tag.err; echo "bench n$N rc=$?"; tail -3 gpurun_out/bench_n${N}_$tag.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_n${N}_$tag.json'))
print('N=$N value %.3e ms/step %.4f' % (d['value'], d['ms_per_step']))
for k,v in (d.get('workloads') or {}).items():
    print(k, {kk:vv for kk,vv in v.items() if kk in ('value','ms_per_step','scaling','n_gpus','max_rel_diff_vs_unsharded','sweep_kernel_ms')})
PY
