#!/bin/bash
# round 2, session b: pipe peaks, ncu captures of the planned sweeps
set -u
mkdir -p gpurun_out
tag=${1:-r2b}
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/peaks_bench scripts/peaks_bench.cu && gpurun_out/peaks_bench > gpurun_out/peaks_$tag.json; cat gpurun_out/peaks_$tag.json
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "tile_shapes" > gpurun_out/pytest_shapes_$tag.log 2>&1; echo "pytest shapes rc=$?"; tail -3 gpurun_out/pytest_shapes_$tag.log
B="--steps 3 --warmup 3 --no-cpu --no-rce --only-main"
export HELIOS_NONISO_MINB=${MINB:-6}
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_sweep_noniso -s 4 -c 2 \
    -o gpurun_out/prof_noniso_$tag -f python bench.py --workload C2 $B > gpurun_out/ncu_noniso_$tag.log 2>&1; echo "ncu noniso rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_sweep_iso -s 4 -c 2 \
    -o gpurun_out/prof_iso_c5_$tag -f python bench.py --workload C5 --batch 32 --steps 3 --warmup 3 > gpurun_out/ncu_iso_$tag.log 2>&1; echo "ncu iso rc=$?"
ls -la gpurun_out/*.ncu-rep
