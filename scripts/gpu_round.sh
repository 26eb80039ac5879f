#!/bin/bash
# One GPU-box session: parity suite, golden vectors from the reference kernels, bench (ours + reference),
# ncu launch list and full captures of the dominant kernels.  Outputs land in gpurun_out/.
set -u
mkdir -p gpurun_out
tag=${1:-r1}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi_$tag.csv 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q -s > gpurun_out/pytest_$tag.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_$tag.log
tail -5 gpurun_out/pytest_$tag.log
timeout 600 python tests/golden/make_ref_kernel_golden.py gpurun_out/ref_kernels_golden.npz > gpurun_out/golden_$tag.log 2>&1; echo "golden rc=$?"
timeout 600 python bench.py --impl reference > gpurun_out/bench_ref_$tag.json 2> gpurun_out/bench_ref_$tag.err; echo "bench ref rc=$?"
timeout 600 python bench.py > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo "bench rc=$?"
timeout 600 python bench.py --workload C1 --no-cpu > gpurun_out/bench_c1_$tag.json 2> gpurun_out/bench_c1_$tag.err
cat gpurun_out/bench_$tag.json
# launch list of the same command (cold-cache, serialised: shares only)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
    --log-file gpurun_out/launches_$tag.csv python bench.py --steps 3 --warmup 3 --no-cpu --no-rce --only-main > gpurun_out/bench_under_ncu_$tag.log 2>&1
echo "ncu list rc=$?"
# full capture of the flux-sweep kernel and the transmission kernel
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_fband -s 4 -c 2 \
    -o gpurun_out/prof_fband_$tag -f python bench.py --steps 3 --warmup 3 --no-cpu --no-rce --only-main > gpurun_out/ncu_fband_$tag.log 2>&1
echo "ncu fband rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_calc_trans|k_pt_gather|k_fdir|k_band_integrate|k_planck_interpol" -c 10 \
    -o gpurun_out/prof_rebuild_$tag -f python bench.py --steps 3 --warmup 3 --no-cpu --no-rce --only-main > gpurun_out/ncu_rebuild_$tag.log 2>&1
echo "ncu rebuild rc=$?"
ls -la gpurun_out | tail -30
