import os, sys, json, numpy as np
sys.path.insert(0, '/root/repo')
import bench
from helios_b200 import backend, runtime
ctx = runtime.set_default_context(backend.Context(0))
flush = ctx.zeros(256 * 1024 * 1024 // 8)
def timeit(comp, q, n=15):
    for _ in range(3):
        comp.populate_spectral_flux_iteratively(q)
    ctx.synchronize()
    ts = []
    for _ in range(n):
        flush.fill_zero()
        e0, e1 = ctx.event(), ctx.event()
        e0.record(); comp.populate_spectral_flux_iteratively(q); e1.record(); e1.synchronize()
        ts.append(e0.time_till(e1))
    return np.median(ts) * 1e3, min(ts) * 1e3
for workload in ("C2", "C1"):
    q, comp = bench._prepare(workload, ctx)
    for ncols in ("8", "16"):
        os.environ["HELIOS_WP_NCOLS"] = ncols
        print(workload, "ncols", ncols, "median %.1f us  min %.1f us" % timeit(comp, q))
    ctx.set_fband_mode(1)
    print(workload, "column-serial median %.1f us" % timeit(comp, q, 3)[0])
    ctx.set_fband_mode(0)
