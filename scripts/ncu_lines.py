#!/usr/bin/env python
"""Per-source-line stall summary of an .ncu-rep (needs -lineinfo and --import-source on).
usage: ncu_lines.py report.ncu-rep [top_n]"""
import collections
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
STALLS = ["stall_barrier", "stall_long_sb", "stall_short_sb", "stall_math", "stall_wait", "stall_mio", "stall_lg",
          "stall_not_selected", "stall_selected", "stall_dispatch", "stall_branch_resolving", "stall_no_inst"]
per_fn = collections.OrderedDict()
fn = f = hdr = None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        f = r[1].split("/")[-1]
        continue
    if r[0] == "Function Name":
        fn = r[1].split("(")[0]
        continue
    if r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or len(r) != len(hdr) or r[0] == "":
        continue
    try:
        s = int(r[hdr.index("# Samples")])
    except ValueError:
        continue
    d = per_fn.setdefault(fn, {})
    key = (f, int(r[0]), r[1].strip()[:80])
    e = d.setdefault(key, [0, collections.Counter()])
    e[0] += s
    for k in STALLS:
        try:
            e[1][k[6:]] += int(r[hdr.index(k)])
        except ValueError:
            pass
for fn, d in per_fn.items():
    tot = sum(e[0] for e in d.values()) or 1
    print("== %s: %d samples" % (fn, tot))
    agg = collections.Counter()
    for e in d.values():
        agg.update(e[1])
    print("   stalls overall: " + " ".join("%s=%.1f%%" % (k, 100.0 * v / tot) for k, v in agg.most_common(8)))
    for (f, l, text), e in sorted(d.items(), key=lambda kv: -kv[1][0])[:top]:
        print("  %5.1f%% %-16s %4d  %-80s %s" % (100.0 * e[0] / tot, f, l, text,
                                                 " ".join("%s=%d" % kv for kv in e[1].most_common(3))))
