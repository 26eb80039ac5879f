#!/bin/bash
# experiment: which part of the planned non-isothermal sweep costs what.  Needs the -DHELIOS_ABLATE build of the library
# (gpurun_in/libhelios_b200_ablate.so); swaps it in on the (scratch) GPU-box copy of the repo only.
# usage: exp_ablate.sh TAG "ab ab ..." "ctas_per_sm ..."
set -u
mkdir -p gpurun_out
out=gpurun_out/ablate_${1:-r2}.txt
: > $out
echo "== product library" >> $out
NPASS=4 timeout 300 python scripts/exp_npass.py C2 C1 >> $out 2>&1
cp gpurun_in/libhelios_b200_ablate.so helios_b200/csrc/libhelios_b200.so
for c in ${3:-0}; do
for ab in ${2:-0 1 2 4 8 16 32 64 65 97 101 125 127}; do
  echo "== ctas_per_sm $c ablate $ab" >> $out
  if [ $c = 0 ]; then HELIOS_SWEEP_ABLATE=$ab NPASS=4 timeout 300 python scripts/exp_npass.py C2 >> $out 2>&1
  else HELIOS_SWEEP_CTAS_PER_SM=$c HELIOS_SWEEP_ABLATE=$ab NPASS=4 timeout 300 python scripts/exp_npass.py C2 >> $out 2>&1; fi
done
done
cat $out
