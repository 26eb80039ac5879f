#!/bin/bash
# bench session: default line (with extras), reference arm, C5 and C4 stand-alone
set -u
mkdir -p gpurun_out
tag=${1:-b}
timeout 1200 python bench.py > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_$tag.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_$tag.json"))
print("value %.3e  ms %.4f  e2e %.3e (%.3f ms)  frac %.3f" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["roofline"]["frac"]))
print("rce", d.get("rce"))
for k,v in (d.get("workloads") or {}).items():
    print(k, {kk:(vv if not isinstance(vv,dict) else {a:b for a,b in vv.items() if a in ("frac","achieved","kernel_ms","value","ms_per_step")}) for kk,vv in v.items() if kk!="workload"})
PY
