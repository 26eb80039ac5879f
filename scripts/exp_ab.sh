#!/bin/bash
# A/B of library builds on ONE box: for every gpurun_in/libhelios_b200_<name>.so named on the command line (and the
# in-tree build as "tree"), the planned sweeps of C2 (non-isothermal, 4 passes), C1 (isothermal) and C4 (isothermal, 1e5
# columns, 1 pass), three rounds interleaved.
set -u
mkdir -p gpurun_out
tag=$1; shift
out=gpurun_out/ab_$tag.txt
: > $out
cp helios_b200/csrc/libhelios_b200.so /tmp/lib_tree.so
for round in 1 2 3; do
  for name in tree "$@"; do
    if [ $name = tree ]; then cp /tmp/lib_tree.so helios_b200/csrc/libhelios_b200.so
    else cp gpurun_in/libhelios_b200_$name.so helios_b200/csrc/libhelios_b200.so; fi
    echo "== $name round $round" >> $out
    NPASS=4 timeout 300 python scripts/exp_npass.py C2 C1 >> $out 2>&1
    NPASS=1 timeout 300 python scripts/exp_npass.py C4 >> $out 2>&1
  done
done
cp /tmp/lib_tree.so helios_b200/csrc/libhelios_b200.so
cat $out
