import os, sys, numpy as np
sys.path.insert(0, '/root/repo')
import bench
from helios_b200 import backend, runtime
ctx = runtime.set_default_context(backend.Context(0))
flush = ctx.zeros(256 * 1024 * 1024 // 8)
MODE = "memset"
def do_flush():
    if MODE == "memset": flush.fill_zero()
    elif MODE == "memset+read": ctx.call("l2_flush", 1)
def timeit(comp, q, n=15):
    for _ in range(3):
        comp.populate_spectral_flux_iteratively(q)
    ctx.synchronize()
    ts = []
    for _ in range(n):
        do_flush()
        e0, e1 = ctx.event(), ctx.event()
        e0.record(); comp.populate_spectral_flux_iteratively(q); e1.record(); e1.synchronize()
        ts.append(e0.time_till(e1))
    return np.median(ts) * 1e3
for workload in ("C2", "C1"):
    q, comp = bench._prepare(workload, ctx)
    for MODE in ("memset", "memset+read", "none"):
        os.environ["HELIOS_FBAND_DBG"] = "0"
        print(workload, "flush", MODE, "full %.1f us" % timeit(comp, q), flush=True)
        os.environ["HELIOS_FBAND_DBG"] = "11"
        print(workload, "flush", MODE, "neither phase, no prefetch %.1f us" % timeit(comp, q), flush=True)
    MODE = "memset+read"
    for dbg, what in ((0, "full"), (1, "no prefetch"), (2, "phase A only (+prefetch)"), (3, "phase A only, no prefetch"),
                      (6, "loads only + prefetch"), (7, "loads only, no prefetch"), (8, "phase B only + prefetch"), (9, "phase B only"), (10, "neither: loop+barriers+prefetch"), (11, "neither, no prefetch")):
        os.environ["HELIOS_FBAND_DBG"] = str(dbg)
        print("%s dbg=%2d %-34s %.1f us" % (workload, dbg, what, timeit(comp, q)), flush=True)
    os.environ["HELIOS_FBAND_DBG"] = "0"
