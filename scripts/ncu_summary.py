#!/usr/bin/env python
"""Key metrics of every launch in an .ncu-rep as a small csv (the files under profiles/r*_ncu_full_*.csv).
usage: ncu_summary.py report.ncu-rep out.csv"""
import csv, io, subprocess, sys

METRICS = """gpu__time_duration.sum dram__bytes_read.sum dram__bytes_write.sum
gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed sm__throughput.avg.pct_of_peak_sustained_elapsed
sm__warps_active.avg.pct_of_peak_sustained_active sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active
smsp__issue_active.avg.pct_of_peak_sustained_active smsp__inst_executed.sum launch__registers_per_thread
launch__grid_size launch__block_size launch__shared_mem_per_block_dynamic launch__occupancy_limit_shared_mem
launch__occupancy_limit_registers lts__t_sector_hit_rate.pct l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum
sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active""".split()

raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
head, units = rows[0], rows[1]
idx = [head.index(m) for m in METRICS]
kname = head.index("Kernel Name")
with open(sys.argv[2], "w", newline="") as f:
    w = csv.writer(f)
    w.writerow(["Kernel Name"] + METRICS)
    w.writerow([""] + [units[i] for i in idx])
    for r in rows[2:]:
        name = r[kname].split("(")[0]
        w.writerow([name] + [r[i] for i in idx])
