#!/bin/bash
set -u
mkdir -p gpurun_out
tag=${1:-r2d}
timeout 300 python scripts/exp_npass.py C2 C1 2>&1 | tee gpurun_out/exp_npass_$tag.log
B="--steps 3 --warmup 3 --no-cpu --no-rce --only-main"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_sweep_noniso -s 4 -c 2 \
    -o gpurun_out/prof_noniso_$tag -f python bench.py --workload C2 $B > gpurun_out/ncu_noniso_$tag.log 2>&1; echo "ncu noniso rc=$?"
