#!/bin/bash
set -u
mkdir -p gpurun_out
tag=${1:-r2e}
for ab in 0 16 17 18 20 24 48 22 23 31 63; do
  echo "== ablate $ab"
  HELIOS_SWEEP_ABLATE=$ab timeout 120 python scripts/exp_npass.py C2 2>&1 | grep -E "npass= (1|4)" 
done | tee gpurun_out/exp_ablate_$tag.log
