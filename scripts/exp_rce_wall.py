#!/usr/bin/env python
"""where the wall clock of the device-loop convergence leg goes (bench.rce_leg, device_loop mode)"""
import os, sys, time, gc
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from helios_b200 import backend, runtime, synthetic
from helios_b200.batch import make_batch

ctx = runtime.set_default_context(backend.Context(0))
for rep in range(3):
    qb, bcomp = make_batch([synthetic.make_store("C2", ctx=ctx, seed=synthetic.SEED)], ctx)
    bcomp.construct_planck_table(qb)
    bcomp.correct_incident_energy(qb)
    bench._quiesce(ctx)
    t = [time.perf_counter()]
    bcomp.radiation_loop(qb)
    t.append(time.perf_counter())
    it = int(qb.converged_at[0])
    t.append(time.perf_counter())
    wall = dict(bcomp.stats.get("radiation_loop_wall", {}))
    del qb, bcomp
    t.append(time.perf_counter())
    ctx.synchronize()
    t.append(time.perf_counter())
    gc.enable()
    print("rep %d: loop %.4f s, read %.4f, del %.4f, sync %.4f | %d iterations | %s" %
          (rep, t[1] - t[0], t[2] - t[1], t[3] - t[2], t[4] - t[3], it, wall), flush=True)
