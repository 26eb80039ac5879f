#!/usr/bin/env python
"""experiment: planned-sweep kernel time as a function of the number of fused passes (slope = per-pass cost,
intercept = staging / lifting / store overhead per launch)"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from helios_b200 import backend, runtime

ctx = runtime.set_default_context(backend.Context(0))
for wl in sys.argv[1:] or ["C2", "C1"]:
    q, comp = bench._prepare(wl, ctx)
    assert q._flux_plan_valid
    for n in [int(v) for v in os.environ.get('NPASS', '1,4,16').split(',')]:
        def run():
            if q.iso == 1:
                ctx.call("fband_iso_planned", q.dev_F_down_wg, q.dev_F_up_wg, q.dev_fband_plan, q.dev_planckband_lay,
                         q.dev_surf_albedo, q.R_star, q.a, q.ninterface, q.nbin, q.f_factor, q.ny, q.dir_beam, n)
            else:
                ctx.call("fband_noniso_planned", q.dev_F_down_wg, q.dev_F_up_wg, q.dev_Fc_down_wg, q.dev_Fc_up_wg,
                         q.dev_fband_plan, q.dev_planckband_lay, q.dev_planckband_int, q.dev_surf_albedo, q.R_star, q.a,
                         q.ninterface, q.nbin, q.f_factor, q.ny, q.dir_beam, n)
        # the event timer of some boxes ticks in 1-2 us steps: time K flushed launches against K flushes alone
        K = 10
        def timed(with_run):
            e0, e1 = ctx.event(), ctx.event()
            e0.record()
            for _ in range(K):
                ctx.call("l2_flush", 1)
                if with_run:
                    run()
            e1.record(); e1.synchronize()
            return e0.time_till(e1) / K
        for _ in range(2):
            timed(True)
        ts = [timed(True) - timed(False) for _ in range(7)]
        print("%s npass=%2d  median %.2f us  min %.2f us" % (wl, n, 1e3 * float(np.median(ts)), 1e3 * min(ts)), flush=True)
