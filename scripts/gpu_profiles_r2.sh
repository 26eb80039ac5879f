#!/bin/bash
# round-2 evidence: launch list of the bench command, full captures of the dominant kernels, sanitizer runs
set -u
mkdir -p gpurun_out
tag=${1:-r2}
B="--steps 3 --warmup 3 --no-cpu --no-rce --only-main"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/launches_$tag.csv python bench.py $B > gpurun_out/bench_under_ncu_$tag.log 2>&1; echo "ncu list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_sweep_noniso -s 4 -c 2 \
    -o gpurun_out/prof_sweep_c2_$tag -f python bench.py --workload C2 $B > gpurun_out/ncu_a_$tag.log 2>&1; echo "ncu C2 rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_sweep_iso -s 4 -c 2 \
    -o gpurun_out/prof_sweep_c5_batch32_$tag -f python bench.py --workload C5 --batch 32 --steps 3 --warmup 3 > gpurun_out/ncu_b_$tag.log 2>&1; echo "ncu C5 rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_sweep_iso -s 4 -c 2 \
    -o gpurun_out/prof_sweep_c4_$tag -f python bench.py --workload C4 --steps 3 --warmup 3 > gpurun_out/ncu_c_$tag.log 2>&1; echo "ncu C4 rc=$?"
timeout 900 ncu --set full --clock-control none -k regex:"k_plan_build|k_calc_trans|k_fdir|k_band_integrate|k_pt_gather|k_iter_prep|k_temp_iter" -c 14 \
    -o gpurun_out/prof_rebuild_$tag -f python bench.py --workload C2 $B > gpurun_out/ncu_d_$tag.log 2>&1; echo "ncu rebuild rc=$?"
timeout 900 ncu --set full --clock-control none -k regex:"k_add_to_mixed_opac" -s 2 -c 2 \
    -o gpurun_out/prof_mixing_$tag -f python bench.py --workload C3 --steps 3 --warmup 3 > gpurun_out/ncu_e_$tag.log 2>&1; echo "ncu mixing rc=$?"
timeout 900 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "tile_shapes and (C1-100 or C2-100 or C1-203 or C2-33) or foreign_plan or tma_staged" > gpurun_out/memcheck_$tag.log 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/memcheck_$tag.log
timeout 900 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "tile_shapes and (C1-100-False or C2-100-False)" > gpurun_out/racecheck_$tag.log 2>&1; echo "racecheck rc=$?"; tail -4 gpurun_out/racecheck_$tag.log
ls -la gpurun_out/*_$tag.ncu-rep
