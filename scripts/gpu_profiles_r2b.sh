#!/bin/bash
# round-2 evidence, final build: launch list, full captures of the C2 / C5 sweeps and of the refresh + integration kernels
set -u
mkdir -p gpurun_out
tag=${1:-r2g}
B="--steps 3 --warmup 3 --no-cpu --no-rce --only-main"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/launches_$tag.csv python bench.py $B > gpurun_out/bench_under_ncu_$tag.log 2>&1; echo "ncu list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_sweep_noniso -s 4 -c 2 \
    -o gpurun_out/prof_sweep_c2_$tag -f python bench.py --workload C2 $B > gpurun_out/ncu_a_$tag.log 2>&1; echo "ncu C2 rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_sweep_iso -s 4 -c 2 \
    -o gpurun_out/prof_sweep_c5_batch32_$tag -f python bench.py --workload C5 --batch 32 --steps 3 --warmup 3 > gpurun_out/ncu_b_$tag.log 2>&1; echo "ncu C5 rc=$?"
timeout 600 ncu --set full --clock-control none -k regex:"k_plan_build|k_calc_trans|k_fdir|k_band_integrate|k_pt_gather|k_iter_prep|k_temp_iter" -c 14 \
    -o gpurun_out/prof_rebuild_$tag -f python bench.py --workload C2 $B > gpurun_out/ncu_d_$tag.log 2>&1; echo "ncu rebuild rc=$?"
ls -la gpurun_out/*_$tag.ncu-rep
