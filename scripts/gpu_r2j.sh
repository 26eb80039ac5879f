#!/bin/bash
set -u
mkdir -p gpurun_out
tag=${1:-r2j}
timeout 900 python -m pytest tests/test_gpu_rce.py tests/test_gpu_refloop.py tests/test_gpu_mixing.py -m gpu -q -s > gpurun_out/pytest_rce_$tag.log 2>&1; echo "pytest rce/refloop/mixing rc=$?"
grep -E "^\[|rce|refloop|passed|failed|Error|error" gpurun_out/pytest_rce_$tag.log | head -40
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
