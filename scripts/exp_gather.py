#!/usr/bin/env python
"""experiment: opac_interpol (premixed table, C2 shape, 207 MB table) with the TMA-staged gather vs the __ldg form"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from helios_b200 import backend, runtime
ctx = runtime.set_default_context(backend.Context(0))
q, comp = bench._prepare("C2", ctx)
ts = []
for k in range(23):
    ctx.call("l2_flush", 1)
    e0, e1 = ctx.event(), ctx.event()
    e0.record(); comp.interpolate_opacities_and_scattering_cross_sections(q); e1.record(); e1.synchronize()
    if k >= 3: ts.append(e0.time_till(e1))
print("HELIOS_PT_GATHER=%s: opac_interpol (lay + int, C2) median %.2f us min %.2f us" % (os.environ.get("HELIOS_PT_GATHER", "default"), 1e3*float(np.median(ts)), 1e3*min(ts)))
