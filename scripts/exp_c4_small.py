#!/usr/bin/env python
"""experiment: the two launches of a C4 flux solve (sweep, band integration) at the per-GPU sizes of a wavelength-sharded
run (1e5 / N bins), on one GPU without the exchange: what the strong-scaling limit of the kernels themselves is"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from helios_b200 import backend, runtime

ctx = runtime.set_default_context(backend.Context(0))
flush = lambda: ctx.call("l2_flush", 1)
for nbin in (100000, 50000, 25000, 12500):
    r = bench.bench_c4(ctx, 0, 1, 20, 5, flush, nbin=nbin, scat=0)
    q, comp = r["q"], r["comp"]
    with ctx.capture() as g:
        comp.populate_spectral_flux_iteratively(q)
        comp.integrate_flux(q)
    def gstep(ev):
        if ev:
            ev[0].record(); ev[1].record()
        g.launch()
        if ev:
            ev[2].record()
    tg, _ = bench._timed(ctx, gstep, 20, 5, flush)
    with ctx.capture() as g1:
        comp.populate_spectral_flux_iteratively(q)
    def g1step(ev):
        if ev:
            ev[0].record(); ev[1].record()
        g1.launch()
        if ev:
            ev[2].record()
    t1, _ = bench._timed(ctx, g1step, 20, 5, flush)
    print("nbin %6d: eager step %.1f us (sweep %.1f), graph step %.1f us, graph sweep only %.1f us -> integration %.1f us"
          % (nbin, 1e3 * r["t_solve"], 1e3 * r["t_fband"], 1e3 * tg, 1e3 * t1, 1e3 * (tg - t1)), flush=True)
