#!/bin/bash
# round 2, first GPU session: the planned sweeps of fband_plan.cu -- parity first, then A/B benches against round 1's kernels
set -u
mkdir -p gpurun_out
tag=${1:-r2a}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi_$tag.csv 2>&1
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "tile_shapes or foreign_plan or unusual" > gpurun_out/pytest_shapes_$tag.log 2>&1; echo "pytest shapes rc=$?"
tail -15 gpurun_out/pytest_shapes_$tag.log
B="--steps 20 --warmup 3 --no-cpu --no-rce --only-main"
for w in C2 C1; do
  timeout 300 python bench.py --workload $w $B > gpurun_out/bench_${w}_$tag.json 2> gpurun_out/bench_${w}_$tag.err; echo "bench $w rc=$?"
  python -c "
import json,sys
d=json.load(open('gpurun_out/bench_${w}_$tag.json'))
print('$w', 'ms/step', d['ms_per_step'], 'kernel_ms', d['roofline']['kernel_ms'], 'frac', round(d['roofline']['frac'],3), 'e2e ms', d['e2e']['ms_per_step'], d['e2e'].get('ms_per_step_with_rebuild'))"
done
for mb in 8; do
  HELIOS_NONISO_CFG=$mb timeout 300 python bench.py --workload C2 $B > gpurun_out/bench_C2_minb${mb}_$tag.json 2> gpurun_out/bench_C2_minb${mb}_$tag.err
  python -c "
import json
d=json.load(open('gpurun_out/bench_C2_minb${mb}_$tag.json'))
print('C2 minb$mb', 'ms/step', d['ms_per_step'], 'kernel_ms', d['roofline']['kernel_ms'])"
done
HELIOS_PLAN_V1=1 timeout 300 python bench.py --workload C2 $B > gpurun_out/bench_C2_v1_$tag.json 2> gpurun_out/bench_C2_v1_$tag.err
python -c "
import json
d=json.load(open('gpurun_out/bench_C2_v1_$tag.json'))
print('C2 v1', 'ms/step', d['ms_per_step'], 'kernel_ms', d['roofline']['kernel_ms'])"
timeout 600 python bench.py --workload C5 --steps 10 --warmup 3 > gpurun_out/bench_C5_$tag.json 2> gpurun_out/bench_C5_$tag.err; echo "bench C5 rc=$?"
python -c "
import json
d=json.load(open('gpurun_out/bench_C5_$tag.json'))
print('C5', 'ms/step', d['ms_per_step'], 'kernel_ms', d['roofline']['kernel_ms'], 'frac', round(d['roofline']['frac'],3))"
for sc in 0 1; do
timeout 600 python bench.py --workload C4 --c4-scat $sc --steps 5 --warmup 3 > gpurun_out/bench_C4s${sc}_$tag.json 2> gpurun_out/bench_C4s${sc}_$tag.err; echo "bench C4 scat$sc rc=$?"
python -c "
import json
d=json.load(open('gpurun_out/bench_C4s${sc}_$tag.json'))
print('C4 scat$sc', 'ms/step', d['ms_per_step'], 'kernel_ms', d['roofline']['kernel_ms'], 'frac', round(d['roofline']['frac'],3))"
done
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_$tag.log 2>&1; echo "pytest full rc=$?"
tail -25 gpurun_out/pytest_$tag.log
