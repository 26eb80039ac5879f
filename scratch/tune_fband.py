import os, sys, json, numpy as np
sys.path.insert(0, '/root/repo')
import bench
from helios_b200 import backend, runtime
ctx = runtime.set_default_context(backend.Context(0))
flush = ctx.zeros(256 * 1024 * 1024 // 8)
for workload in ("C2", "C1"):
    q, comp = bench._prepare(workload, ctx)
    for var in range(0, 8):
        os.environ["HELIOS_CP_VARIANT"] = str(var)
        try:
            for _ in range(3):
                comp.populate_spectral_flux_iteratively(q)
            ctx.synchronize()
        except Exception as e:
            print(workload, "variant", var, "n/a", str(e)[:80]); continue
        ts = []
        for _ in range(15):
            flush.fill_zero()
            e0, e1 = ctx.event(), ctx.event()
            e0.record(); comp.populate_spectral_flux_iteratively(q); e1.record(); e1.synchronize()
            ts.append(e0.time_till(e1))
        print(workload, "variant", var, "median %.1f us  min %.1f us" % (np.median(ts) * 1e3, min(ts) * 1e3))
    ctx.set_fband_mode(1)
    os.environ["HELIOS_CP_VARIANT"] = "0"
    comp.populate_spectral_flux_iteratively(q); ctx.synchronize()
    e0, e1 = ctx.event(), ctx.event(); flush.fill_zero(); e0.record(); comp.populate_spectral_flux_iteratively(q); e1.record(); e1.synchronize()
    print(workload, "column-serial %.1f us" % (e0.time_till(e1) * 1e3))
    ctx.set_fband_mode(0)
