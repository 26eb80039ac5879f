#!/bin/bash
set -u
mkdir -p gpurun_out
tag=${1:-r2f}
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "tile_shapes or foreign_plan or unusual" > gpurun_out/pytest_shapes_$tag.log 2>&1; echo "pytest shapes rc=$?"
tail -12 gpurun_out/pytest_shapes_$tag.log
timeout 300 python scripts/exp_npass.py C2 C1 2>&1 | tee gpurun_out/exp_npass_$tag.log
B="--steps 20 --warmup 3 --no-cpu --no-rce --only-main"
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
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_$tag.log 2>&1; echo "pytest full rc=$?"
tail -8 gpurun_out/pytest_$tag.log
