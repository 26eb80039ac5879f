#!/bin/bash
# N-GPU scaling look (run under gpurun --gpus N): default C2 (weak), C5 (weak), C4 (strong, wavelength-sharded)
N=${1:-8}
mkdir -p gpurun_out
port=29600
for spec in "C2:" "C5:" "C4:--c4-scat 0" "C4:--c4-scat 1"; do
  w=${spec%%:*}; extra=${spec#*:}; port=$((port+1))
  tag=$(echo "${w}_${extra}" | tr -d ' -')
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port \
      bench.py --gpus $N --steps 20 --warmup 3 --workload $w $extra > gpurun_out/scale_${tag}_n$N.json 2> gpurun_out/scale_${tag}_n$N.err
  echo "$w $extra rc=$? lines=$(wc -l < gpurun_out/scale_${tag}_n$N.json)"
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/scale_${tag}_n$N.json").read().strip().splitlines()[-1])
    print("  n_gpus", d["n_gpus"], "value %.3e" % d["value"], "ms_per_step %.4f" % d["ms_per_step"], "scaling", d["scaling"], "rce", (d.get("rce") or {}).get("atmospheres_per_hour"))
except Exception as e:
    print("  parse error", e)
PY
done
