#!/bin/bash
# N GPUs: exchange parity tests (2 ranks), then the default bench at N
set -u
mkdir -p gpurun_out
N=${1:-2}
tag=${2:-r2}
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q -s 2>&1 | grep -E "multi\]|passed|failed|Error" | tail -8
bash scripts/gpu_scale_r2.sh $N $tag
