#!/usr/bin/env python
"""experiment (needs the -DHELIOS_INTEG_TIMING build of the library, gpurun_in/libhelios_b200_integt.so copied over the
in-tree library on the GPU box): SM-cycle stamps of the phases of one block of k_band_integrate on the C2 solve"""
import ctypes, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from helios_b200 import backend, runtime

ctx = runtime.set_default_context(backend.Context(0))
q, comp = bench._prepare("C2", ctx)
lib = backend.lib()
out = (ctypes.c_longlong * 8)()
names = ["start", "after griddepcontrol.wait", "first tile staged", "bins summed", "block sum -> partial", "ticket returned", "end"]
for rep in range(4):
    ctx.call("l2_flush", 1)
    comp.populate_spectral_flux_iteratively(q)
    e0, e1 = ctx.event(), ctx.event()
    e0.record()
    comp.integrate_flux(q)
    e1.record(); e1.synchronize()
    lib.helios_debug_integ_timing(out)
    t = list(out)
    print("rep %d: kernel by events %.1f us; cycles since block start: " % (rep, 1e3 * e0.time_till(e1)) +
          ", ".join("%s %d" % (n, t[k] - t[0]) for k, n in enumerate(names) if k), flush=True)
