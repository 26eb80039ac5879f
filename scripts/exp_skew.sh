#!/bin/bash
# experiment: do the co-resident CTAs of the non-isothermal sweep run faster out of phase? (ablation build)
set -u
mkdir -p gpurun_out
out=gpurun_out/skew_${1:-r2}.txt
: > $out
cp gpurun_in/libhelios_b200_ablate.so helios_b200/csrc/libhelios_b200.so
for sk in 0 500 1000 1500 2000 2500 3000 4000 0; do
  echo "== skew_ns $sk" >> $out
  HELIOS_SWEEP_SKEW_NS=$sk NPASS=4 timeout 300 python scripts/exp_npass.py C2 >> $out 2>&1
done
cat $out
