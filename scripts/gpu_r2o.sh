#!/bin/bash
set -u
mkdir -p gpurun_out
tag=${1:-r2o}
timeout 1200 python -m pytest tests/test_gpu_fullsize.py -m gpu -q -x --durations=8 > gpurun_out/pytest_full_$tag.log 2>&1; echo "pytest fullsize rc=$?"; tail -25 gpurun_out/pytest_full_$tag.log
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "smoothing" > gpurun_out/pytest_smooth_$tag.log 2>&1; echo "pytest smoothing rc=$?"; tail -15 gpurun_out/pytest_smooth_$tag.log
