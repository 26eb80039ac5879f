#!/usr/bin/env python
"""bench.py -- flux-solve throughput of the HELIOS RT hot path on B200 (one JSON line on rank 0).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload C2|C1|C4|C5]

Workload at N=1: BASELINE.json configs[1] ("C2"): 100 layers x 385 bins x 20 Gauss points,
non-isothermal layers, scattering iteration (3*scat+1 = 4 passes), one cloud deck, non-gray albedo and a
direct beam, on seeded synthetic tables (helios_b200/synthetic.py; the Zenodo inputs are not available
offline).  A "step" is one flux solve of the RT iteration: `populate_spectral_flux_iteratively` (all
passes) + `integrate_flux` (computation.py:881-888).

  value      layer*lambda*g-points/s = nlayer*nbin*ny*n_pass / t, inputs resident in HBM, CUDA-event timed per
             step on the launching stream, L2 flushed between steps (helios_l2_flush: 2x L2 overwritten, then
             read back), max over ranks
  e2e        the same metric for one RT iteration through the public API (`Compute.*`) with host buffers, on the
             reference's own schedule (computation.py:851-984): every iteration takes the temperature profile from
             pinned host memory (H2D), rebuilds the Planck terms, runs the flux solve and the temperature step and
             reads the per-interface fluxes + new profile + convergence flags back (D2H); every 10th iteration
             also rebuilds opacities, transmission functions and the direct beam
  rce        converged RCE atmospheres per hour (the second half of BASELINE.json's metric): one atmosphere from
             the isothermal start to the reference's convergence criterion, wall clock
  workloads  (default run only) the other BASELINE.json configurations with the same timing rules:
             C5 = a batched grid of 128 independent atmospheres per GPU, C4 = a 1e5-bin sampling spectrum
  N > 1      C1/C2/C5: every rank owns its own atmosphere(s) (batched grids shard by atmosphere, no
             data-path collective) -> weak scaling;  C4: the spectrum is sharded by wavelength with one fused
             NVLink peer-memory all-reduce of the flux totals per step -> strong scaling

--impl reference runs the reference's own kernels.cu (compiled verbatim to oracle/_ref/, launched with the
block/grid shapes and per-launch device syncs of computation.py) on ONE B200: the reference has no CPU
path, so this is "the reference run on the same box" of BASELINE.json:north_star.  The NumPy oracle timed
on the host cores is reported as `cpu_baseline`.
"""
import argparse
import gc
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "flux_solve_layer_lambda_g_points_per_s"
UNIT = "points/s"


def _peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def _traffic_from_profile(workload, nbatch=None):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the flux-sweep kernel, from the committed
    `ncu --set full` summary of this workload (profiles/); None if no capture of this workload is committed.
    A capture of a batched launch (file name ..._batchN.csv) is scaled to the batch size of this run."""
    import csv
    import glob
    import re
    best = None
    # newest round last (r2_... sorts after r1_...): the latest committed capture of this workload wins
    for path in sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_ncu_full_fband*%s*.csv" % workload.lower()))):
        rows = list(csv.reader(open(path)))
        h, units = rows[0], rows[1]
        vals = []
        for r in rows[2:]:
            if "fband" not in r[0] and "k_sweep" not in r[0]:
                continue
            tot = 0.0
            for key in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                i = h.index(key)
                scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[units[i]]
                tot += float(r[i]) * scale
            vals.append(tot)
        if vals:
            per_launch = sum(vals) / len(vals)
            m = re.search(r"_batch(\d+)\.csv$", path)
            if m and nbatch:
                per_launch *= float(nbatch) / float(m.group(1))
            best = {"bytes_per_launch": per_launch, "source": os.path.relpath(path, ROOT)}
    return best


class ClockSampler(object):
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.file = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(device), "--query-gpu=" + self.QUERY,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=self.file,
                                         stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.proc is None:
            return out
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.file.flush()
        rows = [r.split(",") for r in open(self.file.name).read().strip().splitlines() if r.count(",") >= 8]
        os.unlink(self.file.name)
        if not rows:
            return out
        sm = [float(r[1]) for r in rows if r[1].strip().replace(".", "").isdigit()]
        mx = [float(r[2]) for r in rows if r[2].strip().replace(".", "").isdigit()]
        reasons = set()
        for r in rows:
            for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7), ("sw_power_cap", 8)):
                if "Active" in r[col] and "Not" not in r[col]:
                    reasons.add(name)
        out["sm_mhz"] = float(np.median(sm)) if sm else None
        out["sm_max_mhz"] = max(mx) if mx else None
        out["reasons"] = sorted(reasons)
        out["samples"] = len(rows)
        return out


def _dist():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return world, rank, local


def _prepare(workload, ctx, seed_offset=0):
    """a Store with every coefficient of the flux solve resident on the device"""
    from helios_b200 import synthetic
    from helios_b200.computation import Compute
    q = synthetic.make_store(workload, ctx=ctx, seed=synthetic.SEED + seed_offset)
    n = int(q.nlayer)
    # a realistic, non-isothermal hot-Jupiter profile (deep 2300 K -> 1100 K aloft)
    q.T_lay = np.concatenate([2300.0 - 1200.0 * (np.arange(n) / (n - 1.0)) ** 1.5, [2350.0]])
    synthetic.upload(q)
    comp = Compute(ctx, verbose=False)
    comp.construct_planck_table(q)
    comp.correct_incident_energy(q)
    q.iter_value = np.int32(0)
    refresh(comp, q)
    ctx.synchronize()
    return q, comp


def refresh(comp, q):
    """everything that depends on the temperature profile (computation.py:856-879)"""
    comp.interpolate_temperatures(q)
    comp.interpolate_planck(q)
    comp.interpolate_opacities_and_scattering_cross_sections(q)
    comp.interpolate_meanmolmass(q)
    if q.clouds == 1:
        comp.calc_total_g_0_of_gas_and_clouds(q)
    comp.calculate_transmission(q)
    comp.calculate_delta_z(q)
    q.delta_z_lay = q.dev_delta_z_lay.get()
    comp.hsfunc.calculate_height_z(q)
    q.dev_z_lay.set(q.z_lay)
    comp.calculate_direct_beamflux(q)
    comp.build_flux_plan(q)  # non-isothermal layers: the sweep plan belongs to the refresh (Compute._refresh_atmosphere)


def _no_beam(q):
    """the direct-beam arrays are known to be all zero (fdir_* ran with dir_beam == 0): the sweeps skip them"""
    return int(q.dir_beam) == 0


def _bytes_per_cell(q):
    """bytes ONE fused flux solve has to move per layer*lambda*g cell with the kernel that actually runs (DESIGN.md
    'roofline'): what is loaded once and stored once, nothing that is skipped.
      planned isothermal sweep      3 plan constants (5 with a beam) + previous F_up in, F_down + F_up out, Planck / ny
      planned non-isothermal sweep  8 plan constants (12 with a beam) + previous F_up, Fc_up in, 4 fluxes out,
                                    layer + interface Planck / ny
      unplanned sweeps              the reference's coefficient arrays, minus F_dir and G+- when the beam is
                                    known to be zero"""
    ny = float(q.ny)
    planned = bool(getattr(q, "_flux_plan_valid", False))
    nobeam = _no_beam(q)
    if q.iso == 1:
        if planned:
            return (3 if nobeam else 5) * 8 + 8 + 16 + 8 / ny
        return (4 if nobeam else 6) * 8 + (0 if nobeam else 8) + 8 + 16 + (8 + (8 if q.clouds == 1 else 0)) / ny
    if planned:
        return (8 if nobeam else 12) * 8 + 16 + 32 + 16 / ny
    return (10 if nobeam else 14) * 8 + (0 if nobeam else 16) + 16 + 32 + (2 * 8 + 2 * 8 + (2 * 8 if q.clouds == 1 else 0)) / ny


def _survey_bytes_per_cell(q):
    """SURVEY.md 8(d)'s per-unit figure for the reference's algorithm (all coefficient arrays of fband_* read once):
    80.4 B isothermal, 176 B (+ band terms) non-isothermal.  Reported beside `frac` for continuity with round 1."""
    ny = float(q.ny)
    if q.iso == 1:
        return 6 * 8 + 8 + 8 + 16 + (8 + (8 if q.clouds == 1 else 0)) / ny
    return 14 * 8 + 16 + 16 + 32 + (2 * 8 + 2 * 8 + (2 * 8 if q.clouds == 1 else 0)) / ny


def _sweep_kernel_name(q):
    planned = bool(getattr(q, "_flux_plan_valid", False))
    if q.iso == 1:
        return "k_sweep_iso (planned, TMA-staged)" if planned else "k_fband_wp (iso)"
    return "k_sweep_noniso (planned, TMA-staged)" if planned else "k_fband_wp (noniso)"


# ------------------------------------------------------------------------------------------------------------
# the other BASELINE.json configurations, measured with the same rules (CUDA events on the launching stream,
# L2 flushed between steps, >= 3 warm-up steps).  They ride along in the default line under "workloads".
# ------------------------------------------------------------------------------------------------------------
def _timed(ctx, step, steps, warmup, flush, split=None):
    """average ms per step; `split` (optional) is called between the two halves of a step and returns nothing --
    when given, the time of the first half is returned as well"""
    for _ in range(max(warmup, 3)):
        flush()
        step(None)
    ctx.synchronize()
    ev = [[ctx.event() for _ in range(3)] for _ in range(steps)]
    for k in range(steps):
        flush()
        step(ev[k])
    ctx.synchronize()
    total = sum(e[0].time_till(e[2]) for e in ev) / steps
    first = sum(e[0].time_till(e[1]) for e in ev) / steps
    return total, first


def bench_batch(ctx, rank, world, nbatch, steps, warmup, flush, config="C1"):
    """C5: a grid of independent atmospheres, `nbatch` per GPU, one launch per kernel (helios_b200/batch.py)"""
    from helios_b200 import synthetic, sharding
    from helios_b200.batch import make_batch
    params = sharding.partition_atmospheres(synthetic.grid_parameters(), rank, world)[:nbatch]
    while len(params) < nbatch:  # fewer ranks than the grid was cut for: recycle parameter sets
        params = (params + params)[:nbatch]
    stores = synthetic.make_grid_stores(params, config=config, ctx=ctx)
    n = int(stores[0].nlayer)
    for k, q in enumerate(stores):
        q.T_lay = np.concatenate([2300.0 - 1200.0 * (np.arange(n) / (n - 1.0)) ** 1.5, [2350.0]]) + 3.0 * (k % 16)
    qb, comp = make_batch(stores, ctx)
    del stores
    comp.construct_planck_table(qb)
    comp.correct_incident_energy(qb)
    qb.enter()
    qb.iter_value = np.int32(0)
    comp.interpolate_temperatures(qb)
    comp.interpolate_planck(qb)
    comp._refresh_atmosphere(qb)
    ctx.synchronize()
    npass = comp.n_scat_passes(qb)
    cells = int(qb.nlayer) * int(qb.nbin) * int(qb.ny) * nbatch

    def step(ev):
        if ev:
            ev[0].record()
        comp.populate_spectral_flux_iteratively(qb)
        if ev:
            ev[1].record()
        comp.integrate_flux(qb)
        if ev:
            ev[2].record()

    t_solve, t_fband = _timed(ctx, step, steps, warmup, flush)

    # one full batched RT iteration through the public API with host buffers
    nl1 = int(qb.nlayer) + 1
    T_host = backend_mod().PinnedArray(nbatch * nl1)
    T_host.array[:] = qb.dev_T_lay.get()
    res_host = backend_mod().PinnedArray(nbatch * nl1)
    sums_host = backend_mod().PinnedArray(nbatch, np.int32)
    e0, e1 = ctx.event(), ctx.event()

    def iteration():
        T_host.h2d_async(ctx, qb.dev_T_lay)
        comp.interpolate_temperatures(qb)
        comp.interpolate_planck(qb)
        comp._refresh_atmosphere(qb)
        step(None)
        comp.rad_temp_iteration(qb)
        ctx.call("abort_sum", qb.dev_abort, nl1, qb.dev_abort_sums)
        res_host.d2h_async(ctx, qb.dev_T_lay)
        sums_host.d2h_async(ctx, qb.dev_abort_sums)
        ctx.synchronize()

    for _ in range(2):
        iteration()
    t_e2e = 0.0
    n_e2e = max(3, steps // 2)
    for _ in range(n_e2e):
        flush()
        e0.record()
        iteration()
        e1.record()
        e1.synchronize()
        t_e2e += e0.time_till(e1)
    t_e2e /= n_e2e
    qb.leave()
    bpc = _bytes_per_cell(qb)
    return dict(qb=qb, comp=comp, npass=npass, cells=cells, points=cells * npass, t_solve=t_solve, t_fband=t_fband,
                t_e2e=t_e2e, bpc=bpc, sbpc=_survey_bytes_per_cell(qb), kernel=_sweep_kernel_name(qb), h2d=T_host.nbytes, d2h=res_host.nbytes + sums_host.nbytes,
                workload="C5: %d independent atmospheres per GPU (T_star x log g x opacity scaling grid), each %d layers x "
                         "%d bins x %d gauss points, %s layers, %d fused flux passes; one launch per kernel for the "
                         "whole batch, %d opacity tables shared" % (nbatch, qb.nlayer, qb.nbin, qb.ny,
                                                                    "isothermal" if qb.iso == 1 else "non-isothermal",
                                                                    npass, qb.ntables))


def backend_mod():
    from helios_b200 import backend
    return backend


def bench_c4(ctx, rank, world, steps, warmup, flush, nbin=100000, scat=1, dim=8000, step_k=2, reuse=None):
    """C4: post-processing of a high-resolution opacity-sampling spectrum (1e5 bins x 1 point), wavelength-sharded
    across the ranks with the fused NVLink all-reduce of the flux totals (helios_b200/sharding.py)"""
    from helios_b200 import synthetic, sharding
    from helios_b200.computation import Compute
    if reuse is not None:
        q, comp = reuse["q"], reuse["comp"]
        q.scat = np.int32(scat)
    else:
        q = synthetic.make_store("C4", ctx=ctx, nbin=nbin, plancktable_dim=dim, plancktable_step=step_k)
        q.scat = np.int32(scat)
        n = int(q.nlayer)
        q.T_lay = np.concatenate([2300.0 - 1200.0 * (np.arange(n) / (n - 1.0)) ** 1.5, [2350.0]])
        if world > 1:
            sharding.shard_store(q, rank, world)
        synthetic.upload(q)
        if world > 1:
            sharding.attach_flux_allreduce(q, ctx, rank, world)
        comp = Compute(ctx, verbose=False)
        comp.construct_planck_table(q)
    q.iter_value = np.int32(0)
    refresh(comp, q)
    ctx.synchronize()
    npass = comp.n_scat_passes(q)
    cells = int(q.nlayer) * int(q.nbin) * int(q.ny)

    def step(ev):
        if ev:
            ev[0].record()
        comp.populate_spectral_flux_iteratively(q)
        if ev:
            ev[1].record()
        comp.integrate_flux(q)
        if ev:
            ev[2].record()

    t_solve, t_fband = _timed(ctx, step, steps, warmup, flush)
    return dict(q=q, comp=comp, npass=npass, cells=cells, points=cells * npass, t_solve=t_solve, t_fband=t_fband,
                bpc=_bytes_per_cell(q), sbpc=_survey_bytes_per_cell(q), kernel=_sweep_kernel_name(q),
                workload="C4: post-processing spectrum, %d layers x %d bins x 1 point (this rank: %d bins), %d fused "
                         "flux passes (scat=%d), wavelength-sharded over %d GPU(s)" %
                         (q.nlayer, nbin, q.nbin, npass, scat, world))


def bench_mixing(ctx, flush, reps=5):
    """C3: on-the-fly mixing of 10 species (BASELINE.json configs[2]) -- what every 10th RT iteration does instead of
    the premixed-table gather: per species a (P,T) interpolation of its own k-table plus correlated-k summation or
    random overlap (400 k-sums sorted and rebinned per cell).  The species tables stay resident in HBM."""
    from helios_b200 import synthetic, host
    from helios_b200.computation import Compute
    out = {}
    q = synthetic.make_store("C3", ctx=ctx, kcoeff_mixing="RO")
    n = int(q.nlayer)
    q.T_lay = np.concatenate([2300.0 - 1200.0 * (np.arange(n) / (n - 1.0)) ** 1.5, [2350.0]])
    synthetic.upload(q)
    comp = Compute(ctx, verbose=False)
    comp.interpolate_temperatures(q)
    host.calculate_meanmolecularmass(q)
    cells = int(q.nlayer) * int(q.nbin)
    nspec = len(q.species_list)
    for mixing in ("RO", "correlated-k"):
        q.kcoeff_mixing = mixing
        times = []
        for k in range(reps + 2):
            host.nullify_opac_scat_arrays(q)
            flush()
            e0, e1 = ctx.event(), ctx.event()
            e0.record()
            comp.calculate_total_opacity_and_scat_cross_sections_from_species(q)
            e1.record()
            e1.synchronize()
            if k >= 2:
                times.append(e0.time_till(e1))
        ms = float(np.median(times))
        out[mixing] = {"ms_per_species_loop": ms, "species": nspec, "cells_x_i": cells,
                       "mixed_cells_per_s": cells * nspec / (ms * 1e-3),
                       "k_combinations_sorted_per_s": (cells * (nspec - 1) * 400 / (ms * 1e-3)) if mixing == "RO" else None}
    out["workload"] = ("C3: %d species, %d layers x %d bins x %d gauss points, isothermal layers; one full species loop "
                       "(interpolation + mixing + scattering cross sections), tables resident in HBM" %
                       (nspec, q.nlayer, q.nbin, q.ny))
    return out


def _quiesce(ctx):
    """before a wall-clock leg: release what earlier legs left behind now (every DeviceArray that dies frees its
    buffer with a device-synchronising cudaFree) and keep the collector out of the timed region"""
    gc.collect()
    ctx.synchronize()
    gc.disable()


def rce_leg(ctx, workload, seed_offset=0):
    """converged RCE atmospheres per hour: one atmosphere from the standard isothermal start to the reference's own
    convergence criterion, wall clock with all host logic included.  Two drivers of the same kernels:
      host_loop    Compute.radiation_loop + convection_loop: the reference's loop structure (one poll per iteration)
      device_loop  BatchCompute.radiation_loop with nbatch = 1: iteration counter and convergence latch on the
                   device, blocks of 10 iterations replayed as one CUDA graph; bit-identical result
                   (tests/test_gpu_batch.py); the convection loop then runs on the host loop as in the reference"""
    from helios_b200 import synthetic
    from helios_b200.batch import make_batch
    from helios_b200.computation import Compute
    out = {}
    for mode in ("host_loop", "device_loop", "host_loop", "device_loop"):  # wall clock on a shared box: best of two
        q = synthetic.make_store(workload, ctx=ctx, seed=synthetic.SEED + seed_offset)
        status = "converged"
        conv_iters = 0
        if mode == "host_loop":
            synthetic.upload(q)
            comp = Compute(ctx, verbose=False)
            comp.construct_planck_table(q)
            comp.correct_incident_energy(q)
            _quiesce(ctx)
            t0 = time.perf_counter()
            try:
                comp.radiation_loop(q, None, None, None)
                rad_iters = int(q.iter_value)
                if q.convection == 1:
                    comp.convection_loop(q, None, None, None)
                    conv_iters = int(q.iter_value)
            except SystemExit:
                status, rad_iters = "iteration limit", int(q.iter_value)
        else:
            single = synthetic.upload(q)
            qb, bcomp = make_batch([synthetic.make_store(workload, ctx=ctx, seed=synthetic.SEED + seed_offset)], ctx)
            comp = Compute(ctx, verbose=False)
            bcomp.construct_planck_table(qb)
            bcomp.correct_incident_energy(qb)
            comp.construct_planck_table(single)
            comp.correct_incident_energy(single)
            _quiesce(ctx)
            t0 = time.perf_counter()
            try:
                bcomp.radiation_loop(qb)
                rad_iters = int(qb.converged_at[0])
                if q.convection == 1:
                    # hand the converged state over to the host-driven convection loop (C:992-1174)
                    for name in ("T_lay", "F_net", "F_up_tot", "F_down_tot", "abort", "T_store", "delta_t_prefactor"):
                        getattr(single, "dev_" + name).copy_from(getattr(qb, "dev_" + name))
                    single.marked_red = (single.dev_abort.get() == 0).astype(np.float64)
                    comp.convection_loop(single, None, None, None)
                    conv_iters = int(single.iter_value)
            except SystemExit:
                status, rad_iters = "iteration limit", int(qb.iter_value)
            del qb, bcomp
        ctx.synchronize()
        dt = time.perf_counter() - t0
        gc.enable()
        if mode not in out or dt < out[mode]["seconds"]:
            out[mode] = {"seconds": dt, "radiation_iterations": rad_iters, "convection_iterations": conv_iters,
                         "status": status, "ms_per_iteration": 1e3 * dt / max(1, rad_iters + conv_iters)}
    return out


def rce_batch(ctx, nbatch=32):
    """the same for a grid of C1 atmospheres advanced together (one launch per kernel for the whole batch)"""
    from helios_b200 import synthetic, sharding
    from helios_b200.batch import make_batch
    params = sharding.partition_atmospheres(synthetic.grid_parameters(), 0, 1024 // nbatch)
    qb, comp = make_batch(synthetic.make_grid_stores(params, config="C1", ctx=ctx), ctx)
    comp.construct_planck_table(qb)
    comp.correct_incident_energy(qb)
    _quiesce(ctx)
    t0 = time.perf_counter()
    comp.radiation_loop(qb)
    ctx.synchronize()
    dt = time.perf_counter() - t0
    gc.enable()
    at = [int(v) for v in qb.converged_at]
    return {"atmospheres": nbatch, "seconds": dt, "atmospheres_per_hour": nbatch * 3600.0 / dt,
            "iterations_to_convergence_min_max": [min(at), max(at)], "iterations_run": int(qb.iter_value),
            "what": "BatchCompute.radiation_loop: %d C1 atmospheres of the T_star x log g x opacity grid advanced together "
                    "to the reference's convergence criterion, converged ones frozen by the on-device latch" % nbatch}


def _roofline(kernel, bpc, cells, t_kernel_ms, npass, workload, nbatch=None, survey_bpc=None):
    peak, peak_src = _peaks()
    traffic = _traffic_from_profile(workload, nbatch)
    t_k = t_kernel_ms * 1e-3
    achieved = bpc * cells / t_k / 1e9
    out = {"bound": "hbm", "kernel": kernel, "achieved": achieved, "peak": peak, "unit": "GB/s",
           "frac": achieved / peak, "traffic": (traffic or {}).get("bytes_per_launch"),
           "traffic_source": (traffic or {}).get("source"), "algorithmic_bytes_per_launch": bpc * cells,
           "peak_source": peak_src, "bytes_per_cell_per_solve": bpc, "kernel_ms": t_kernel_ms,
           "bytes_basis": "what the running kernel must load and store once per flux solve (skipped arrays not counted)"}
    if out["traffic"]:
        out["traffic_over_algorithmic"] = out["traffic"] / (bpc * cells)
    if survey_bpc is not None:
        out["frac_at_survey_bytes"] = survey_bpc * cells / t_k / 1e9 / peak
        out["survey_bytes_per_cell"] = survey_bpc
    if npass >= 8:
        # long pass sequences run from registers: the launch is bound by the dependent fp64 chain, not by HBM
        out["note"] = "compute-bound launch (%d fused passes in registers): the HBM fraction is not the binding roofline" % npass
    return out


def run_ours(args):
    import torch
    import torch.distributed as dist
    from helios_b200 import backend, runtime
    world, rank, local = _dist()
    torch.cuda.set_device(local)
    if world > 1:
        # NCCL writes its version / debug lines to stdout by default; stdout carries the JSON line only
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = runtime.set_default_context(backend.Context(local))
    def flush():
        ctx.call("l2_flush", 1)

    steps, warmup = args.steps, max(args.warmup, 3)

    def barrier():
        ctx.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def reduce_max(values):
        t = torch.tensor(values, dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return [float(v) for v in t.tolist()]

    sampler = ClockSampler(local) if rank == 0 else None
    base = {"metric": METRIC, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warmup,
            "higher_is_better": True, "vs_baseline": None, "dtype": "f64", "data": "synthetic"}
    l2 = ("flushed between timed steps, outside the event bracket: a 256 MiB buffer (2x L2) is overwritten and then "
          "read back, so the timed kernels start on a cold L2 that holds no dirty lines of the flusher")

    if args.workload == "C5":
        barrier()
        r = bench_batch(ctx, rank, world, args.batch, steps, warmup, flush)
        barrier()
        t_solve, t_fband, t_e2e = reduce_max([r["t_solve"], r["t_fband"], r["t_e2e"]])
        line = dict(base, value=world * r["points"] / (t_solve * 1e-3), ms_per_step=t_solve, scaling="weak",
                    config={"workload": r["workload"], "l2": l2, "sharding": "atmospheres dealt round-robin to the ranks, no collective"},
                    e2e={"value": world * r["points"] / (t_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": r["h2d"],
                         "d2h_bytes_per_step": r["d2h"], "ms_per_step": t_e2e,
                         "what": "one full batched RT iteration via BatchCompute.* (T profiles from pinned host, rebuild, "
                                 "flux solve, temperature step, profiles + convergence sums to host)"},
                    gpu_launches=None,
                    roofline=_roofline("%s, all %d passes fused, %d atmospheres" % (r["kernel"], r["npass"], args.batch),
                                       r["bpc"], r["cells"], t_fband, r["npass"], "C5", args.batch, r["sbpc"]))
    elif args.workload == "C4":
        barrier()
        r = bench_c4(ctx, rank, world, steps, warmup, flush, scat=args.c4_scat)
        barrier()
        t_solve, t_fband = reduce_max([r["t_solve"], r["t_fband"]])
        pts = reduce_max([float(r["points"])])  # ranks hold nbin/world +- 1 bins
        total_points = 100 * 100000 * r["npass"]
        line = dict(base, value=total_points / (t_solve * 1e-3), ms_per_step=t_solve, scaling="strong",
                    config={"workload": r["workload"], "l2": l2,
                            "sharding": "contiguous wavelength ranges per rank; one fused NVLink peer-memory all-reduce of "
                                        "the per-interface flux totals per step"},
                    e2e=None, gpu_launches=None,
                    roofline=_roofline("%s, %d passes fused" % (r["kernel"], r["npass"]), r["bpc"], r["cells"], t_fband,
                                       r["npass"], "C4", None, r["sbpc"]))
        del pts
    else:
        line = _run_single(args, ctx, flush, base, l2, world, rank, barrier, reduce_max)
        if rank == 0 and world == 1 and not args.only_main:
            extra = {}
            try:
                t0 = time.perf_counter()
                r = bench_batch(ctx, 0, 8, args.batch, max(5, steps // 5), 3, flush)
                extra["C5_batched_grid"] = {
                    "workload": r["workload"], "value": r["points"] / (r["t_solve"] * 1e-3), "unit": UNIT,
                    "ms_per_step": r["t_solve"],
                    "e2e": {"value": r["points"] / (r["t_e2e"] * 1e-3), "unit": UNIT, "ms_per_step": r["t_e2e"],
                            "h2d_bytes_per_step": r["h2d"], "d2h_bytes_per_step": r["d2h"]},
                    "roofline": _roofline("%s, %d passes fused, %d atmospheres" % (r["kernel"], r["npass"], args.batch),
                                          r["bpc"], r["cells"], r["t_fband"], r["npass"], "C5", args.batch, r["sbpc"]),
                    "setup_s": time.perf_counter() - t0}
                del r
                extra["C5_batched_grid"]["rce"] = rce_batch(ctx, 32)
                # the headline physics (C2: non-isothermal, clouds, beam; planned sweep) as a batch of 32 atmospheres:
                # the dominant kernel without the single-atmosphere tail effect (1925 tiles on 592 CTAs)
                r = bench_batch(ctx, 0, 32, 32, max(5, steps // 5), 3, flush, config="C2")
                extra["C2_batched_grid_32"] = {
                    "workload": r["workload"].replace("C5:", "C2 physics:"), "value": r["points"] / (r["t_solve"] * 1e-3),
                    "unit": UNIT, "ms_per_step": r["t_solve"],
                    "roofline": _roofline("%s, %d passes fused, 32 atmospheres" % (r["kernel"], r["npass"]),
                                          r["bpc"], r["cells"], r["t_fband"], r["npass"], "C2batch", 32, r["sbpc"])}
                del r
            except Exception as e:  # noqa: BLE001 -- the main line must survive a failing extra
                extra["C5_batched_grid"] = {"error": repr(e)}
            r = None
            for scat in (0, 1):
                key = "C4_spectrum_1e5_bins_scat%d" % scat
                try:
                    r = bench_c4(ctx, 0, 1, max(3, steps // 10), 3, flush, scat=scat, reuse=r)
                    extra[key] = {"workload": r["workload"], "value": r["points"] / (r["t_solve"] * 1e-3), "unit": UNIT,
                                  "ms_per_step": r["t_solve"],
                                  "roofline": _roofline("%s, %d passes fused" % (r["kernel"], r["npass"]), r["bpc"], r["cells"],
                                                        r["t_fband"], r["npass"], "C4", None, r["sbpc"])}
                except Exception as e:  # noqa: BLE001
                    extra[key] = {"error": repr(e)}
            try:
                extra["C3_on_the_fly_mixing"] = bench_mixing(ctx, flush)
            except Exception as e:  # noqa: BLE001
                extra["C3_on_the_fly_mixing"] = {"error": repr(e)}
            line["workloads"] = extra
    clocks = sampler.stop() if sampler else None
    if rank == 0:
        line["clocks"] = clocks
        emit(line)
    if world > 1:
        dist.destroy_process_group()


def _run_single(args, ctx, flush, base, l2, world, rank, barrier, reduce_max):
    """C1 / C2: one atmosphere per rank (weak scaling, no collective)"""
    from helios_b200 import backend
    q, comp = _prepare(args.workload, ctx, seed_offset=rank)
    npass = comp.n_scat_passes(q)
    cells = int(q.nlayer) * int(q.nbin) * int(q.ny)
    points = cells * npass

    def flux_solve(events=None):
        if events:
            events[0].record()
        comp.populate_spectral_flux_iteratively(q)
        if events:
            events[1].record()
        comp.integrate_flux(q)
        if events:
            events[2].record()

    for _ in range(max(args.warmup, 3)):
        flush()
        flux_solve()
    barrier()
    ev = [[ctx.event() for _ in range(3)] for _ in range(args.steps)]
    launches0 = ctx.launch_count()
    barrier()
    for k in range(args.steps):
        flush()  # L2 flush between timed steps (not inside the event bracket)
        flux_solve(ev[k])
    barrier()
    launches = ctx.launch_count() - launches0
    t_solve = sum(e[0].time_till(e[2]) for e in ev)
    t_fband = sum(e[0].time_till(e[1]) for e in ev)

    # ---- end to end through the public API, host buffers, pinned staging
    T_host = backend.PinnedArray(int(q.nlayer) + 1)
    T_host.array[:] = q.dev_T_lay.get()
    nint = int(q.ninterface)
    res_host = backend.PinnedArray(3 * nint + int(q.nlayer) + 1)
    abort_host = backend.PinnedArray(int(q.nlayer) + 1, np.int32)
    d2h = res_host.nbytes + abort_host.nbytes
    h2d = T_host.nbytes
    res_dev = ctx.zeros(3 * nint + int(q.nlayer) + 1)

    def iteration(k):
        """iteration k of the radiation loop as the reference schedules it (C:851-984): the temperature-dependent
        opacities, transmission functions and direct beam are rebuilt every 10th iteration (C:860), the Planck
        terms, the flux solve and the temperature step run every iteration"""
        T_host.h2d_async(ctx, q.dev_T_lay)
        if k % 10 == 0:
            refresh(comp, q)
        else:
            comp.prepare_iteration(q)  # interpolate_temperatures + interpolate_planck (C:856-857), one launch
        flux_solve()
        comp.rad_temp_iteration(q)
        for j, name in enumerate(("F_net", "F_up_tot", "F_down_tot")):
            res_dev.view(j * nint, nint).copy_from(getattr(q, "dev_" + name))
        res_dev.view(3 * nint, int(q.nlayer) + 1).copy_from(q.dev_T_lay)
        res_host.d2h_async(ctx, res_dev)
        abort_host.d2h_async(ctx, q.dev_abort)
        ctx.synchronize()

    for k in range(10):
        iteration(k)
    barrier()
    e0, e1 = ctx.event(), ctx.event()
    e2e_steps = 10 * max(2, args.steps // 10)  # whole blocks of the reference's 10-iteration schedule
    t_e2e = t_e2e_refresh = 0.0
    for k in range(e2e_steps):
        flush()
        e0.record()
        iteration(k)
        e1.record()
        e1.synchronize()
        dt = e0.time_till(e1)
        t_e2e += dt
        if k % 10 == 0:
            t_e2e_refresh += dt
    barrier()
    rce = None
    if not args.no_rce:
        legs = rce_leg(ctx, args.workload, seed_offset=rank)
        barrier()
        rce_s = reduce_max([legs["device_loop"]["seconds"]])[0]
        rce = dict(legs, atmospheres_per_hour=world * 3600.0 / rce_s, seconds_max_over_ranks=rce_s,
                   what="one %s atmosphere per GPU from the isothermal start to the reference's convergence criterion "
                        "(rad_convergence_limit 1e-8): radiation loop on the device (CUDA-graph blocks of 10 iterations) + "
                        "convection loop, wall clock incl. host logic; host_loop = the reference's loop structure; each the "
                        "faster of two runs" % args.workload)
    t_solve, t_fband, t_e2e = reduce_max([t_solve, t_fband, t_e2e])
    line = dict(base, value=world * points * args.steps / (t_solve * 1e-3), ms_per_step=t_solve / args.steps,
                scaling="weak",
                config={"workload": "%s: %d layers x %d bins x %d gauss points, %s layers, %d fused flux passes, "
                                    "clouds=%d, dir_beam=%d; one atmosphere per GPU%s" %
                                    (args.workload, q.nlayer, q.nbin, q.ny, "isothermal" if q.iso == 1 else "non-isothermal",
                                     npass, q.clouds, q.dir_beam,
                                     "; planned sweep (the Planck-independent step constants are formed with the "
                                     "coefficients at the opacity refresh, every 10th iteration as in C:860, and are "
                                     "inputs of the flux solve)" if getattr(q, "_flux_plan_valid", False) else ""),
                        "l2": l2, "sharding": "one atmosphere per rank, no collective"},
                e2e={"value": world * points * e2e_steps / (t_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d,
                     "d2h_bytes_per_step": d2h, "ms_per_step": t_e2e / e2e_steps,
                     "ms_per_step_with_rebuild": t_e2e_refresh / (e2e_steps // 10), "steps": e2e_steps,
                     "what": "one RT iteration through Compute.* with host buffers, averaged over whole blocks of the "
                             "reference's schedule (C:851-984): every iteration = T profile H2D from pinned memory, Planck "
                             "terms, flux solve (all passes), band integration, temperature step, fluxes + profile + "
                             "convergence flags D2H; every 10th iteration additionally rebuilds opacities, transmission "
                             "functions and the direct beam (C:860)"},
                gpu_launches=int(launches),
                roofline=_roofline("%s, all %d passes fused" % (_sweep_kernel_name(q), npass),
                                   _bytes_per_cell(q), cells, t_fband / args.steps, npass, args.workload, None,
                                   _survey_bytes_per_cell(q)),
                rce=rce)
    if rank == 0 and world == 1 and not args.no_cpu:
        line["cpu_baseline"] = cpu_baseline(q, args.workload)
    return line


def cpu_baseline(q_dev, workload):
    """the NumPy oracle's flux solve on this box's host cores (single process: NumPy ufuncs are single-threaded)"""
    from oracle.pipeline import HostMirror, OracleCompute
    m = HostMirror(q_dev)
    oc = OracleCompute()
    npass = (3 if m.singlewalk == 0 else 1000) * int(m.scat) + 1
    points = int(m.nlayer) * int(m.nbin) * int(m.ny) * npass
    reps, t0 = 0, time.perf_counter()
    while True:
        oc.populate_spectral_flux_iteratively(m)
        oc.integrate_flux(m)
        reps += 1
        if time.perf_counter() - t0 > 10.0 or reps >= 20:
            break
    dt = time.perf_counter() - t0
    return {"value": points * reps / dt, "unit": UNIT, "cores": 1, "kind": "port",
            "sample": "%d full %s flux solves (fband x%d + integrate_flux), NumPy fp64 oracle, %.1f s" % (reps, workload, npass, dt)}


def run_reference(args):
    """the reference's kernels.cu on one B200, launch shapes + per-launch syncs of computation.py"""
    world, rank, local = _dist()
    if rank != 0:
        return
    from oracle import ref_gpu
    if not ref_gpu.available():
        emit({"impl": "reference", "unavailable": "oracle/_ref/helios_ref.cubin was not built (needs /root/reference at build time)"})
        return
    from helios_b200 import backend, runtime
    ctx = runtime.set_default_context(backend.Context(local))
    q, comp = _prepare(args.workload, ctx)
    ref = ref_gpu.RefCompute(local)
    npass = comp.n_scat_passes(q)
    points = int(q.nlayer) * int(q.nbin) * int(q.ny) * npass
    def flush():
        ctx.call("l2_flush", 1)

    sampler = ClockSampler(local)

    def flux_solve():
        ref.populate_spectral_flux_iteratively(q)
        ref.integrate_flux(q)

    for _ in range(max(args.warmup, 3)):
        flux_solve()
    t = 0.0
    n0 = ref.mod.launches
    for _ in range(args.steps):
        flush()
        ctx.synchronize()
        t0 = time.perf_counter()  # the reference syncs the device after every launch, so wall clock == device time
        flux_solve()
        t += time.perf_counter() - t0
    clocks = sampler.stop()
    value = points * args.steps / t
    emit({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": 1, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": t / args.steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "%s: %d layers x %d bins x %d gauss points, %d fband launches + integrate_flux_double, "
                               "reference kernels.cu on one B200 (the reference is single-GPU, no CPU path)" %
                               (args.workload, q.nlayer, q.nbin, q.ny, npass),
                   "l2": "flushed between timed steps"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": 0, "kind": "reference",
                         "sample": "reference kernels.cu (verbatim cubin) on GPU 0; %d launches per step" % ((ref.mod.launches - n0) // args.steps)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": int(ref.mod.launches - n0), "clocks": clocks})


_REAL_STDOUT = None


def _quiet_stdout():
    """Libraries (NCCL prints its version line) write to fd 1; the contract is ONE JSON line on stdout.  Point
    fd 1 at stderr for the duration of the run and keep the real stdout for the result line."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line):
    out = _REAL_STDOUT if _REAL_STDOUT is not None else sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="C2", choices=["C1", "C2", "C4", "C5"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-rce", action="store_true", help="skip the converged-atmospheres-per-hour leg")
    ap.add_argument("--only-main", action="store_true", help="skip the extra workloads (C5 batch, C4 spectrum)")
    ap.add_argument("--batch", type=int, default=128, help="atmospheres per GPU of the C5 workload")
    ap.add_argument("--c4-scat", type=int, default=0, help="C4: 1 = 1001 scattering passes as the reference's "
                                                            "post-processing does, 0 = single pass")
    args = ap.parse_args()
    _quiet_stdout()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
