#!/bin/bash
set -u
mkdir -p gpurun_out
tag=${1:-r2r}
timeout 900 python -m pytest tests/test_gpu_convect.py tests/test_gpu_fullsize.py tests/test_gpu_rce.py tests/test_gpu_batch.py -m gpu -q -s 2>&1 | grep -E "^\[|convect|passed|failed|Error|error|assert" | head -40
