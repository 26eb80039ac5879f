#!/bin/bash
# one full capture (with source) of the C2 sweep kernel
set -u
mkdir -p gpurun_out
tag=${1:-x}
B="--steps 3 --warmup 3 --no-cpu --no-rce --only-main"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_sweep_noniso -s 4 -c 1 \
    -o gpurun_out/prof_sweep_c2_$tag -f python bench.py --workload C2 $B > gpurun_out/ncu_a_$tag.log 2>&1; echo "ncu C2 rc=$?"
ls -la gpurun_out/prof_sweep_c2_$tag.ncu-rep
