#!/bin/bash
# parity of the sweeps (tile shapes, column counts, full size) + the sweep timings (C2, C1, C4)
set -u
mkdir -p gpurun_out
tag=${1:-q}
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_batch.py -m gpu -q -x > gpurun_out/pytest_$tag.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_$tag.log
NPASS=4 timeout 300 python scripts/exp_npass.py C2 C1
NPASS=1 timeout 300 python scripts/exp_npass.py C4
