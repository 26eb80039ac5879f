#!/bin/bash
# end-of-round validation: smoke, full GPU suite, reference arm + default bench (timed), ncu launch list + full captures
set -u
mkdir -p gpurun_out
tag=${1:-final}
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_$tag.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_$tag.log
t0=$(date +%s); timeout 900 python bench.py --impl reference > gpurun_out/bench_ref_$tag.json 2> gpurun_out/bench_ref_$tag.err; echo "bench ref rc=$? $(( $(date +%s) - t0 ))s"
t0=$(date +%s); timeout 1200 python bench.py > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo "bench rc=$? $(( $(date +%s) - t0 ))s"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv \
    --log-file gpurun_out/launches_$tag.csv python bench.py --steps 3 --warmup 3 --no-cpu --no-rce --only-main > gpurun_out/bench_under_ncu_$tag.log 2>&1
echo "ncu list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_fband -s 4 -c 2 \
    -o gpurun_out/prof_fband_$tag -f python bench.py --steps 3 --warmup 3 --no-cpu --no-rce --only-main > gpurun_out/ncu_fband_$tag.log 2>&1
echo "ncu fband rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_calc_trans|k_pt_gather|k_fdir|k_band_integrate|k_plan_build|k_planck_interpol|k_temp_iter" -c 14 \
    -o gpurun_out/prof_rebuild_$tag -f python bench.py --steps 3 --warmup 3 --no-cpu --no-rce --only-main > gpurun_out/ncu_rebuild_$tag.log 2>&1
echo "ncu rebuild rc=$?"
