import os, sys, numpy as np
sys.path.insert(0, '/root/repo')
import bench
from helios_b200 import backend, runtime
ctx = runtime.set_default_context(backend.Context(0))
flush = ctx.zeros(256 * 1024 * 1024 // 8)
for workload in ("C2", "C1"):
    q, comp = bench._prepare(workload, ctx)
    for scat, sw in ((1, 0), (0, 0), (1, 1)):
        q.scat = np.int32(scat); q.singlewalk = np.int32(sw)
        npass = comp.n_scat_passes(q)
        for flushit in (True, False):
            for _ in range(3): comp.populate_spectral_flux_iteratively(q)
            ctx.synchronize()
            ts = []
            for _ in range(10):
                if flushit: flush.fill_zero()
                e0, e1 = ctx.event(), ctx.event()
                e0.record(); comp.populate_spectral_flux_iteratively(q); e1.record(); e1.synchronize()
                ts.append(e0.time_till(e1))
            print(workload, "npass", npass, "flush" if flushit else "warmL2", "median %.1f us" % (np.median(ts) * 1e3))
