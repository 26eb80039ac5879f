#!/bin/bash
set -u
mkdir -p gpurun_out
tag=${1:-r2h}
B="--steps 3 --warmup 3 --no-cpu --no-rce --only-main"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_sweep_noniso -s 4 -c 2 \
    -o gpurun_out/prof_noniso_$tag -f python bench.py --workload C2 $B > gpurun_out/ncu_noniso_$tag.log 2>&1; echo "ncu noniso rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_sweep_iso -s 4 -c 2 \
    -o gpurun_out/prof_iso_c5_$tag -f python bench.py --workload C5 --batch 32 --steps 3 --warmup 3 > gpurun_out/ncu_iso_$tag.log 2>&1; echo "ncu iso rc=$?"
