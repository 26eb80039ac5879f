#!/usr/bin/env python
"""experiment: which synthetic internal temperature gives a run whose radiative solution is convectively unstable and
whose radiative-convective loop converges (for the RCE leg of bench.py)"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from helios_b200 import backend, runtime, synthetic, host
from helios_b200.computation import Compute

ctx = runtime.set_default_context(backend.Context(0))
small = dict(nbin=37, nlayer=24, ntemp=12, npress=8, plancktable_dim=700, plancktable_step=10)
for kw, label in ((small, "small"), ({}, "full")):
    for T_intern in (100.0, 200.0, 300.0, 500.0):
        q = synthetic.make_store("C2", ctx=ctx, **kw)
        q.mu_star = np.float64(np.cos((180 - 50.0) * np.pi / 180.0))
        q.T_intern = np.float64(T_intern)
        host.calc_F_intern(q)
        q.max_nr_iterations = 30000
        synthetic.upload(q)
        comp = Compute(ctx, verbose=False)
        comp.construct_planck_table(q)
        comp.correct_incident_energy(q)
        t0 = time.perf_counter()
        status = "ok"
        rad = conv = 0
        try:
            comp.radiation_loop(q, None, None, None)
            rad = int(q.iter_value)
            comp.convection_loop(q, None, None, None)
            conv = int(q.iter_value)
        except SystemExit:
            status = "iteration limit"
        ctx.synchronize()
        print("%s T_intern=%5.0f: %s rad %d conv %d, convective layers %d, %.2f s" %
              (label, T_intern, status, rad, conv, int(np.sum(q.conv_layer)) if q.conv_layer is not None else -1,
               time.perf_counter() - t0), flush=True)
