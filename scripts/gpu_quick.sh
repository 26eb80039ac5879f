#!/bin/bash
# quick GPU session: parity suite, golden vectors, default bench
set -u
mkdir -p gpurun_out
tag=${1:-q}
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/pytest_$tag.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_$tag.log
grep -E "passed|failed|\[rce\]" gpurun_out/pytest_$tag.log | tail -8
timeout 600 python tests/golden/make_ref_kernel_golden.py gpurun_out/ref_kernels_golden.npz > gpurun_out/golden_$tag.log 2>&1; echo "golden rc=$?"; tail -3 gpurun_out/golden_$tag.log
timeout 600 python bench.py > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo "bench rc=$?"
cat gpurun_out/bench_$tag.json
