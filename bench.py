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


def _fp64_peaks():
    """measured pipe peaks of this pool's B200 (scripts/peaks_bench.cu -> profiles/r2_measured_fp64_peaks.json)"""
    path = os.path.join(ROOT, "profiles", "r2_measured_fp64_peaks.json")
    if os.path.exists(path):
        return json.load(open(path)), "measured (profiles/r2_measured_fp64_peaks.json, scripts/peaks_bench.cu)"
    return {"fp64_fma_tflops": 36.4, "dfma_per_clk_per_sm": 62.5, "ddiv_rn_gops": 1245.0, "exp_f64_gops": 875.0,
            "sqrt_f64_gops": 1222.0, "sms": 148, "sm_clock_ghz_nominal": 1.965}, "fallback (round-2 measurement)"


def _fp64_roofline(kernel, ops, t_ms, what):
    """compute-bound launches: `ops` = {"fma": n, "div": n, "exp": n, "sqrt": n, "cmp": n} algorithmic fp64 operations per
    launch; the bound is the time the measured pipe rates need for them, frac = bound / measured time"""
    pk, src = _fp64_peaks()
    fma_rate = pk["fp64_fma_tflops"] * 0.5e12  # FMA (or MUL / ADD / compare) instructions per second
    t_bound = (ops.get("fma", 0) + ops.get("cmp", 0)) / fma_rate + ops.get("div", 0) / (pk["ddiv_rn_gops"] * 1e9) + \
        ops.get("exp", 0) / (pk["exp_f64_gops"] * 1e9) + ops.get("sqrt", 0) / (pk["sqrt_f64_gops"] * 1e9)
    flops = 2.0 * ops.get("fma", 0) + ops.get("cmp", 0) + ops.get("div", 0) + ops.get("exp", 0) + ops.get("sqrt", 0)
    return {"bound": "fp64", "kernel": kernel, "achieved": flops / (t_ms * 1e-3) / 1e12, "peak": pk["fp64_fma_tflops"],
            "unit": "TFLOP/s", "frac": t_bound / (t_ms * 1e-3), "traffic": None, "kernel_ms": t_ms, "operations": ops,
            "peak_source": src, "what": what}


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
    launch_mode = "eager (two C-ABI calls per step)"
    if world > 1:
        # a sharded flux solve is two launches (sweep, integration + fused exchange): recorded once and replayed as a
        # CUDA graph, so that the host's call overhead does not bound the step at 1e5 / N bins per GPU
        with ctx.capture() as g:
            comp.populate_spectral_flux_iteratively(q)
            comp.integrate_flux(q)

        def gstep(ev):
            if ev:
                ev[0].record()
                ev[1].record()
            g.launch()
            if ev:
                ev[2].record()

        t_solve, _ = _timed(ctx, gstep, steps, warmup, flush)
        launch_mode = "CUDA graph of the two launches"
    return dict(q=q, comp=comp, npass=npass, cells=cells, points=cells * npass, t_solve=t_solve, t_fband=t_fband,
                launch_mode=launch_mode,
                totals=(q.dev_F_up_tot.get(), q.dev_F_down_tot.get()),
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
        if mixing == "RO":
            # per (x, i) cell and species: 400 k-sums + 400 weights, a 512-wide bitonic network = 256 * 45 = 11,520
            # compare-exchanges (one fp64 compare each), a 400-long weight scan and 20 interpolations of ~4 FMA
            per_cell = {"fma": 400 + 400 + 80, "cmp": 11520}
            out[mixing]["roofline"] = _fp64_roofline(
                "k_add_to_mixed_opac (random overlap) + k_pt_gather, %d species" % nspec,
                {k: v * cells * (nspec - 1) for k, v in per_cell.items()}, ms,
                "fp64-pipe operations of the sort network and the rebinning; the loop also interpolates every species' table")
    out["workload"] = ("C3: %d species, %d layers x %d bins x %d gauss points, isothermal layers; one full species loop "
                       "(interpolation + mixing + scattering cross sections), tables resident in HBM" %
                       (nspec, q.nlayer, q.nbin, q.ny))
    return out


def sharded_self_check(ctx, rank, world):
    """N >= 2: the wavelength-sharded flux solve with the fused NVLink exchange against the unsharded solve of the same
    (small) spectrum, both run by every rank; returns the largest relative difference of the per-interface totals"""
    from helios_b200 import backend, synthetic, sharding
    from helios_b200.computation import Compute
    kw = dict(nbin=4096, nlayer=100, ntemp=12, npress=8, plancktable_dim=700, plancktable_step=10)
    out = []
    for sharded in (False, True):
        q = synthetic.make_store("C4", ctx=ctx, **kw)
        q.scat = np.int32(1)
        q.singlewalk = np.int32(0)  # 4 passes
        n = int(q.nlayer)
        q.T_lay = np.concatenate([2300.0 - 1200.0 * (np.arange(n) / (n - 1.0)) ** 1.5, [2350.0]])
        if sharded:
            sharding.shard_store(q, rank, world)
        synthetic.upload(q)
        if sharded:
            sharding.attach_flux_allreduce(q, ctx, rank, world)
        comp = Compute(ctx, verbose=False)
        comp.construct_planck_table(q)
        q.iter_value = np.int32(0)
        refresh(comp, q)
        for _ in range(2):
            comp.populate_spectral_flux_iteratively(q)
            comp.integrate_flux(q)
        out.append((q.dev_F_up_tot.get(), q.dev_F_down_tot.get()))
        ctx.synchronize()
        if sharded:
            backend._check(backend.lib().helios_comm_destroy(ctx.handle), "helios_comm_destroy")
    worst = 0.0
    for a, b in zip(out[0], out[1]):
        worst = max(worst, float(np.max(np.abs(a - b) / np.maximum(np.abs(a), 1e-300))))
    return worst


def multi_gpu_workloads(ctx, rank, world, args, flush, barrier, reduce_max):
    """the two partitioned configurations of BASELINE.json at N > 1, every rank taking part:
      C4  the 1e5-bin spectrum sharded by wavelength (strong scaling), exchange fused into the integration kernel
      C5  the batched grid, 128 atmospheres per GPU (weak scaling, no collective)"""
    from helios_b200 import backend
    steps = max(5, args.steps // 5)
    extra = {}
    # C5 first: the exchanges of the sharded legs enable peer access between the GPUs of this process group, after which
    # device-scope fences (one per block of the band integration: tens of thousands per batched launch) cost visibly more
    # (measured at N=8: batched step 2.2 -> 3.4 ms when run after the sharded legs)
    gc.collect()
    barrier()
    r = bench_batch(ctx, rank, world, args.batch, steps, 3, flush)
    barrier()
    t_solve, t_fband, t_e2e = reduce_max([r["t_solve"], r["t_fband"], r["t_e2e"]])
    extra["C5_batched_grid"] = {"workload": r["workload"], "value": world * r["points"] / (t_solve * 1e-3), "unit": UNIT,
                                "n_gpus": world, "scaling": "weak", "ms_per_step": t_solve, "sweep_kernel_ms": t_fband,
                                "e2e": {"value": world * r["points"] / (t_e2e * 1e-3), "unit": UNIT, "ms_per_step": t_e2e,
                                        "h2d_bytes_per_step": r["h2d"], "d2h_bytes_per_step": r["d2h"]}}
    del r
    gc.collect()
    barrier()
    check = sharded_self_check(ctx, rank, world)
    check = reduce_max([check])[0]
    extra["sharded_self_check"] = {"max_rel_diff_vs_unsharded": check, "what": "4096-bin spectrum, 100 layers, 2 flux solves "
                                   "of 4 passes: per-interface flux totals, wavelength-sharded over %d GPUs with the "
                                   "fused exchange vs unsharded" % world}
    r = None
    for scat in (0, 1):
        key = "C4_spectrum_1e5_bins_scat%d" % scat
        barrier()
        r = bench_c4(ctx, rank, world, max(3, steps // 2) if scat else steps, 3, flush, scat=scat, reuse=r)
        barrier()
        t_solve, t_fband = reduce_max([r["t_solve"], r["t_fband"]])
        total_points = 100 * 100000 * r["npass"]
        extra[key] = {"workload": r["workload"], "value": total_points / (t_solve * 1e-3), "unit": UNIT, "n_gpus": world,
                      "scaling": "strong", "ms_per_step": t_solve, "sweep_kernel_ms": t_fband,
                      "launches_per_step": 2, "launch_mode": r["launch_mode"],
                      "exchange": "fused into k_band_integrate's epilogue: peer stores + flags over NVLink, rank-order sum"}
    backend._check(backend.lib().helios_comm_destroy(ctx.handle), "helios_comm_destroy")
    del r
    return extra


def _sweep_roofline(r, tag, nbatch=None):
    """HBM roofline for a flux solve of a few fused passes; for long pass sequences (post-processing: 1001 passes run from
    registers) the launch is bound by the fp64 pipe instead: 4 multiply-adds per cell and pass (cc and the walk step, both
    sweeps) are the recurrence's own arithmetic"""
    if r["npass"] >= 8:
        return _fp64_roofline("%s, %d passes fused" % (r["kernel"], r["npass"]), {"fma": 4.0 * r["cells"] * r["npass"]},
                              r["t_fband"], "the two-stream recurrence itself: 2 FMA per cell, sweep and pass (source term, "
                              "flux step); the layer-parallel evaluation executes ~3x that (composition + scan)")
    return _roofline("%s, %d passes fused" % (r["kernel"], r["npass"]), r["bpc"], r["cells"], r["t_fband"], r["npass"], tag,
                     nbatch, r["sbpc"])


def bench_rebuild(ctx, flush, workload="C2", reps=7):
    """the kernels of an opacity refresh (every 10th RT iteration, C:860-879) and the one-time Planck table, each timed
    alone with CUDA events on a cold L2; fp64 rooflines for the compute-bound ones from their algorithmic operation counts
    and the measured pipe rates (profiles/r2_measured_fp64_peaks.json)"""
    q, comp = _prepare(workload, ctx)
    cells = int(q.nlayer) * int(q.nbin) * int(q.ny)

    def timed(fn):
        ts = []
        for k in range(reps + 2):
            flush()
            e0, e1 = ctx.event(), ctx.event()
            e0.record()
            fn()
            e1.record()
            e1.synchronize()
            if k >= 2:
                ts.append(e0.time_till(e1))
        return float(np.median(ts))

    out = {"workload": "%s: refresh kernels, %d layers x %d bins x %d gauss points" % (workload, q.nlayer, q.nbin, q.ny)}
    t = timed(lambda: comp.calculate_transmission(q))
    halves = 1 if q.iso == 1 else 2
    # per half-layer cell: 1 exp, 2 sqrt, ~8 divisions, ~50 multiply-adds (trans.cu: cell_coeffs)
    out["calculate_transmission"] = {"ms": t, "roofline": _fp64_roofline(
        "k_calc_trans_%s" % ("iso" if q.iso == 1 else "noniso"),
        {"exp": 1.0 * halves * cells, "sqrt": 2.0 * halves * cells, "div": 8.0 * halves * cells, "fma": 50.0 * halves * cells}, t,
        "1 exp + 2 sqrt + ~8 div + ~50 FMA per half-layer cell; the kernel also stores 8 / 16 arrays (%.0f MB)"
        % (cells * 8 * 8 * halves / 1e6))}
    out["calculate_direct_beamflux"] = {"ms": timed(lambda: comp.calculate_direct_beamflux(q))}
    out["interpolate_opacities"] = {"ms": timed(lambda: comp.interpolate_opacities_and_scattering_cross_sections(q))}
    out["build_flux_plan"] = {"ms": timed(lambda: comp.build_flux_plan(q))}
    t = timed(lambda: comp.construct_planck_table(q))
    entries = (int(q.plancktable_dim) + 1) * int(q.nbin)
    # K:362-416: 199-term series at both bin edges: 398 exp and ~8 multiply-adds per term
    out["construct_planck_table"] = {"ms": t, "roofline": _fp64_roofline(
        "k_plancktable", {"exp": 398.0 * entries, "fma": 398.0 * 8 * entries}, t,
        "one-time: %d table entries x 398 exp (199-term series at both bin edges)" % entries)}
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
        loop_wall = None
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
                loop_wall = dict(bcomp.stats.get("radiation_loop_wall", {}), gpu_ms=bcomp.stats.get("radiation_loop_ms"))
                if q.convection == 1:
                    # hand the converged state over to the host-driven convection loop (C:992-1174)
                    for name in ("T_lay", "F_net", "F_up_tot", "F_down_tot", "abort", "T_store", "delta_t_prefactor"):
                        getattr(single, "dev_" + name).copy_from(getattr(qb, "dev_" + name))
                    single.marked_red = (single.dev_abort.get() == 0).astype(np.float64)
                    comp.convection_loop(single, None, None, None)
                    conv_iters = int(single.iter_value)
            except SystemExit:
                status, rad_iters = "iteration limit", int(qb.iter_value)
        ctx.synchronize()
        dt = time.perf_counter() - t0  # the converged state is on the host; releasing the buffers is not part of the run
        if mode == "device_loop":
            del qb, bcomp
        gc.enable()
        if mode not in out or dt < out[mode]["seconds"]:
            out[mode] = {"seconds": dt, "radiation_iterations": rad_iters, "convection_iterations": conv_iters,
                         "status": status, "ms_per_iteration": 1e3 * dt / max(1, rad_iters + conv_iters)}
            if mode == "device_loop":
                out[mode]["wall_breakdown"] = loop_wall
    return out


def convection_leg(ctx, n_iter=400):
    """The radiative-convective loop (C:992-1174) actually iterating: the C2 atmosphere from a super-adiabatic start
    profile, first `n_iter` iterations (the synthetic opacities give no C2 variant whose convective loop converges --
    with T_intern >= 200 K it runs into the iteration limit with the product's and with the reference's kernels alike --
    so the loop BODY is timed).  host path: convective adjustment, marking and equilibrium test as host functions with
    their ~12 transfers per iteration (the reference's structure); device path: csrc/convect.cu, 16 bytes per iteration."""
    from helios_b200 import synthetic
    from helios_b200.computation import Compute
    out = {}
    for mode in ("host_functions", "device_kernels"):
        q = synthetic.make_store("C2", ctx=ctx)
        p = np.asarray(q.p_lay)
        T = np.maximum(3200.0 * (p / p[0]) ** 0.45, 900.0)  # steeper than the dry adiabat (kappa = 2/7) at depth
        q.T_lay = np.append(T, T[0] * 1.05)
        q.max_nr_iterations = n_iter
        synthetic.upload(q)
        comp = Compute(ctx, verbose=False)
        comp.device_convection = mode == "device_kernels"
        comp.construct_planck_table(q)
        comp.correct_incident_energy(q)
        _quiesce(ctx)
        t0 = time.perf_counter()
        try:
            comp.convection_loop(q, None, None, None)
        except SystemExit:
            pass
        ctx.synchronize()
        dt = time.perf_counter() - t0
        gc.enable()
        out[mode] = {"iterations": int(q.iter_value), "seconds": dt, "ms_per_iteration": 1e3 * dt / max(1, int(q.iter_value)),
                     "convective_layers": int(np.sum(q.conv_layer))}
    out["what"] = convection_leg.__doc__.split("\n")[0].strip()
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


def e2e_leg(ctx, workload, flush, steps, rank, barrier):
    """the metric measured end to end: one RT iteration per step through the public API with HOST buffers.  The
    temperature profile comes from pinned host memory every iteration and the per-interface fluxes, the new profile and
    the convergence flags go back to the host every iteration; the copies are part of the recorded iteration."""
    from helios_b200 import backend, synthetic
    from helios_b200.batch import make_batch
    q1 = synthetic.make_store(workload, ctx=ctx, seed=synthetic.SEED + rank)
    n = int(q1.nlayer)
    q1.T_lay = np.concatenate([2300.0 - 1200.0 * (np.arange(n) / (n - 1.0)) ** 1.5, [2350.0]])
    qb, bc = make_batch([q1], ctx)
    bc.construct_planck_table(qb)
    bc.correct_incident_energy(qb)
    qb.enter()
    lib = backend.lib()
    backend._check(lib.helios_ctx_batch_device_iteration(ctx.handle, 1), "helios_ctx_batch_device_iteration")
    qb.state(reset=True)
    qb.iter_value = np.int32(0)  # the kernels read the device counter; the argument is ignored
    nint = int(qb.ninterface)
    T_host = backend.PinnedArray(n + 1)
    T_host.array[:] = qb.dev_T_lay.get()
    # what the host of the reference's loop looks at (C:927-952): the convergence flags every iteration, the profile
    # and the net flux for its reports
    # The three arrays live side by side in ONE device block (the Store's arrays are re-pointed to windows of it), so
    # that the report of an iteration is one device->host copy instead of three
    words = nint + (n + 1) + (n + 2) // 2  # doubles: F_net | T_lay | abort (int32, rounded up)
    report_dev = ctx.zeros(words)
    report_host = backend.PinnedArray(words)
    views = {"F_net": report_dev.view(0, nint), "T_lay": report_dev.view(nint, n + 1),
             "abort": report_dev.view(nint + n + 1, n + 1, dtype=np.int32)}
    for name, v in views.items():
        v.copy_from(getattr(qb, "dev_" + name))
        setattr(qb, "dev_" + name, v)
    ctx.synchronize()

    def body(refresh):
        T_host.h2d_async(ctx, qb.dev_T_lay)
        bc._iteration(qb, refresh=refresh, heights=False, fused=True)
        report_host.d2h_async(ctx, report_dev)

    for k in range(10):  # eager first block: sizes the library's scratch buffers, builds the plan buffer
        body(k == 0)
    ctx.synchronize()
    graphs = {}
    for refresh in (True, False):
        with ctx.capture() as g:
            body(refresh)
        graphs[refresh] = g
    for k in range(10):
        graphs[k == 0].launch()
    ctx.synchronize()
    barrier()
    e0, e1 = ctx.event(), ctx.event()
    e2e_steps = 10 * max(2, steps // 10)  # whole blocks of the reference's 10-iteration schedule
    t_total = t_refresh = 0.0
    for k in range(e2e_steps):
        flush()
        e0.record()
        graphs[k % 10 == 0].launch()
        e1.record()
        e1.synchronize()
        dt = e0.time_till(e1)
        t_total += dt
        if k % 10 == 0:
            t_refresh += dt
    barrier()
    backend._check(lib.helios_ctx_batch_device_iteration(ctx.handle, 0), "helios_ctx_batch_device_iteration")
    qb.leave()
    # the host's view of the last iteration equals the device's (the copy is the one the timed region made)
    rep = report_host.array
    assert np.array_equal(rep[:nint], qb.dev_F_net.get()) and np.array_equal(rep[nint:nint + n + 1], qb.dev_T_lay.get())
    return {"t_total": t_total, "t_refresh": t_refresh, "steps": e2e_steps, "h2d": T_host.nbytes,
            "d2h": report_host.nbytes, "calls": 1}


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

    if args.workload == "C3":
        m = bench_mixing(ctx, flush)
        ro = m["RO"]
        line = dict(base, metric="mixed_cells_per_s", unit="cells/s", value=ro["mixed_cells_per_s"],
                    ms_per_step=ro["ms_per_species_loop"], scaling="weak", config={"workload": m["workload"]},
                    e2e=None, gpu_launches=None, roofline=ro.get("roofline"), mixing=m)
    elif args.workload == "C5":
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
                    roofline=_sweep_roofline(dict(r, t_fband=t_fband), "C4"))
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
                                  "n_gpus": 1, "scaling": "strong", "ms_per_step": r["t_solve"], "sweep_kernel_ms": r["t_fband"],
                                  "roofline": _sweep_roofline(r, "C4")}
                except Exception as e:  # noqa: BLE001
                    extra[key] = {"error": repr(e)}
            try:
                extra["C3_on_the_fly_mixing"] = bench_mixing(ctx, flush)
            except Exception as e:  # noqa: BLE001
                extra["C3_on_the_fly_mixing"] = {"error": repr(e)}
            try:
                extra["C2_refresh_kernels"] = bench_rebuild(ctx, flush)
            except Exception as e:  # noqa: BLE001
                extra["C2_refresh_kernels"] = {"error": repr(e)}
            line["workloads"] = extra
        elif world > 1 and not args.only_main:
            # every N: the wavelength-sharded spectrum (strong scaling, with the NVLink exchange) and the batched grid
            # (weak scaling) ride along, so that the scaling run measures the partitioned configurations as well
            extra = multi_gpu_workloads(ctx, rank, world, args, flush, barrier, reduce_max)
            if rank == 0:
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
    # the sweep kernel alone, eagerly, with an event between the two launches of the step
    ev = [[ctx.event() for _ in range(3)] for _ in range(args.steps)]
    for k in range(args.steps):
        flush()
        flux_solve(ev[k])
    ctx.synchronize()
    t_fband = sum(e[0].time_till(e[1]) for e in ev)
    # the timed step: the two launches of a flux solve (all fused passes; band integration) recorded once and replayed as
    # a CUDA graph -- with N processes on one host the eager form measures the host's launch jitter between the two
    # launches (0.060 ms at N=1, 0.070 ms at N=8 for identical kernels), the graph measures the GPU
    with ctx.capture() as step_graph:
        flux_solve()
    for _ in range(3):
        flush()
        step_graph.launch()
    barrier()
    ev = [[ctx.event() for _ in range(2)] for _ in range(args.steps)]
    barrier()
    for k in range(args.steps):
        flush()  # L2 flush between timed steps (not inside the event bracket)
        ev[k][0].record()
        step_graph.launch()
        ev[k][1].record()
    barrier()
    launches = 2 * args.steps  # sweep + integration per replayed step
    t_solve = sum(e[0].time_till(e[1]) for e in ev)

    # ---- end to end through the public API, host buffers, pinned staging
    e2e = e2e_leg(ctx, args.workload, flush, args.steps, rank, barrier)
    t_e2e, t_e2e_refresh, e2e_steps, h2d, d2h = (e2e["t_total"], e2e["t_refresh"], e2e["steps"], e2e["h2d"], e2e["d2h"])
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
        if rank == 0 and world == 1:
            try:
                rce["convection_loop"] = convection_leg(ctx)
            except Exception as e:  # noqa: BLE001
                rce["convection_loop"] = {"error": repr(e)}
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
                        "l2": l2, "sharding": "one atmosphere per rank, no collective",
                        "launch": "the step's two launches replayed as one CUDA graph"},
                e2e={"value": world * points * e2e_steps / (t_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d,
                     "d2h_bytes_per_step": d2h, "ms_per_step": t_e2e / e2e_steps,
                     "ms_per_step_with_rebuild": t_e2e_refresh / (e2e_steps // 10), "steps": e2e_steps,
                     "launch_calls_per_iteration": e2e["calls"],
                     "what": "one RT iteration through the public API (BatchCompute with nbatch = 1: iteration counter on "
                             "the device, the iteration recorded once with ctx.capture() and replayed as a CUDA graph) with "
                             "host buffers, averaged over whole blocks of the reference's schedule (C:851-984): every "
                             "iteration = T profile H2D from pinned memory, Planck terms, flux solve (all passes), band "
                             "integration, temperature step + convergence latch, fluxes + profile + convergence flags D2H; "
                             "every 10th iteration additionally rebuilds opacities, transmission functions, the direct beam "
                             "and the sweep plan (C:860).  Same kernels as the eager Compute.* path (bit-identical, "
                             "tests/test_gpu_batch.py), one host call per iteration instead of ~12"},
                gpu_launches=int(launches),
                roofline=_roofline("%s, all %d passes fused" % (_sweep_kernel_name(q), npass),
                                   _bytes_per_cell(q), cells, t_fband / args.steps, npass, args.workload, None,
                                   _survey_bytes_per_cell(q)),
                rce=rce)
    if rank == 0 and world == 1 and not args.no_cpu:
        line["cpu_baseline"] = cpu_baseline(q, args.workload)
    return line


def _cpu_worker(job):
    """one host process of the lambda-sharded CPU baseline: the NumPy oracle's flux solve on this process's share of the
    wavelength bins (NumPy ufuncs are single-threaded, so the host cores are used by sharding the spectrum)"""
    workload, rank, world, seconds = job
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    from helios_b200 import synthetic, sharding
    from oracle.pipeline import mirror_from_host, OracleCompute
    # the Planck table only has to cover the profile: a coarse one keeps the untimed set-up short
    q = synthetic.make_store(workload, plancktable_dim=800, plancktable_step=10)
    n = int(q.nlayer)
    q.T_lay = np.concatenate([2300.0 - 1200.0 * (np.arange(n) / (n - 1.0)) ** 1.5, [2350.0]])
    if world > 1:
        sharding.shard_store(q, rank, world)
    m = mirror_from_host(q)
    oc = OracleCompute()
    m.iter_value = np.int32(0)
    for stage in ("construct_planck_table", "interpolate_temperatures", "interpolate_planck",
                  "interpolate_opacities_and_scattering_cross_sections", "interpolate_meanmolmass"):
        getattr(oc, stage)(m)
    if m.clouds == 1:
        oc.calc_total_g_0_of_gas_and_clouds(m)
    oc.calculate_transmission(m)
    m.dev_z_lay = np.zeros(n)
    oc.calculate_direct_beamflux(m)
    npass = (3 if m.singlewalk == 0 else 1000) * int(m.scat) + 1
    points = int(m.nlayer) * int(m.nbin) * int(m.ny) * npass
    oc.populate_spectral_flux_iteratively(m)  # warm-up
    reps, t0 = 0, time.perf_counter()
    while True:
        oc.populate_spectral_flux_iteratively(m)
        oc.integrate_flux(m)
        reps += 1
        if time.perf_counter() - t0 > seconds:
            break
    return points, reps, time.perf_counter() - t0


def cpu_baseline(q_dev, workload):
    """the NumPy oracle's flux solve on this box's host cores: wavelength-sharded over os.cpu_count() processes (each a
    single-threaded NumPy process on its share of the bins), plus the single-process figure"""
    import multiprocessing as mp
    ncores = max(1, min(os.cpu_count() or 1, int(q_dev.nbin)))
    pts, reps, dt = _cpu_worker((workload, 0, 1, 5.0))
    single = pts * reps / dt
    out = {"value": single, "unit": UNIT, "cores": 1, "kind": "port",
           "sample": "%d full %s flux solves, NumPy fp64 oracle, one process, %.1f s" % (reps, workload, dt)}
    if ncores > 1:
        try:
            with mp.get_context("spawn").Pool(ncores) as pool:
                res = pool.map(_cpu_worker, [(workload, r, ncores, 8.0) for r in range(ncores)])
            total = sum(p * k for p, k, _ in res) / max(t for _, _, t in res)
            out = {"value": total, "unit": UNIT, "cores": ncores, "kind": "port", "single_core_value": single,
                   "sample": "%s flux solve (fband x passes + integrate_flux), NumPy fp64 oracle, wavelength-sharded over %d "
                             "host processes, %d..%d solves of a %d-bin share each in ~8 s; single process: %.3g points/s"
                             % (workload, ncores, min(k for _, k, _ in res), max(k for _, k, _ in res),
                                int(q_dev.nbin) // ncores, single)}
        except Exception as e:  # noqa: BLE001 -- the single-process figure stands
            out["multi_process_error"] = repr(e)
    return out


def _ref_inputs(workload, **kw):
    """host-side inputs of a synthetic run as a plain dict, made in a CHILD process: the reference arm itself never
    imports helios_b200 (whose synthetic-input generator the child uses) and never maps libhelios_b200.so"""
    import pickle
    fd, path = tempfile.mkstemp(suffix=".pkl")
    os.close(fd)
    code = ("import sys; sys.path.insert(0, %r); from oracle.refshim import runner; runner.dump_host_store(%r, %r, **%r)"
            % (ROOT, workload, path, kw))
    subprocess.run([sys.executable, "-c", code], check=True, stdout=subprocess.DEVNULL)
    host = pickle.load(open(path, "rb"))
    os.unlink(path)
    return host


def _reference_mixing(run, driver, local):
    """C3 through the reference's own species loop (C:1454-1501), as it runs every 10th iteration: per species an H2D of
    its full k-table, interpolation, random overlap with a single-thread bubble sort per cell (K:3263)"""
    q, hs = run.q, run.hsfunc
    sampler = ClockSampler(local)
    q.iter_value = np.int32(0)
    run.comp.interpolate_temperatures(q)
    times = []
    for k in range(3):
        hs.calculate_meanmolecularmass(q)
        hs.nullify_opac_scat_arrays(q)
        driver.Context.synchronize()
        t0 = time.perf_counter()
        run.comp.calculate_total_opacity_and_scat_cross_sections_from_species(q)
        driver.Context.synchronize()
        times.append(time.perf_counter() - t0)
    ms = 1e3 * min(times)
    cells, nspec = int(q.nlayer) * int(q.nbin), len(q.species_list)
    emit({"impl": "reference", "metric": "mixed_cells_per_s", "value": cells * nspec / (ms * 1e-3), "unit": "cells/s",
          "n_gpus": 1, "steps": 3, "warmup": 0, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
          "vs_baseline": None, "dtype": "f64", "data": "synthetic",
          "config": {"workload": "C3: %d species, %d layers x %d bins x %d gauss points, %s; one species loop of the reference's "
                                 "computation.py (table H2D + interpolation + mixing per species)" %
                                 (nspec, q.nlayer, q.nbin, q.ny, q.kcoeff_mixing)},
          "clocks": sampler.stop()})


def run_reference(args):
    """The reference's stock code path on ONE B200: its unmodified source/computation.py + quantities.py +
    host_functions.py (byte code in oracle/_ref/helios_py) driving its unmodified kernels.cu (oracle/_ref/helios_ref.cubin)
    over the PyCUDA stand-in of oracle/refshim -- launch shapes, per-launch device syncs and host round trips exactly as
    the reference has them.  The reference has no CPU path, so this is "the reference run on the same box" of
    BASELINE.json:north_star.  This process imports neither helios_b200 nor its shared library."""
    world, rank, local = _dist()
    if rank != 0:
        return
    os.environ.setdefault("REFSHIM_DEVICE", str(local))
    from oracle.refshim import runner
    if not runner.available():
        emit({"impl": "reference", "unavailable": "oracle/_ref (reference cubin + byte-compiled computation.py) was not "
                                                  "built: needs /root/reference at build time (make -C oracle ref refpy)"})
        return
    workload = args.workload if args.workload in ("C1", "C2", "C3") else "C2"
    host = _ref_inputs(workload)
    n = int(host["nlayer"])
    host["T_lay"] = np.concatenate([2300.0 - 1200.0 * (np.arange(n) / (n - 1.0)) ** 1.5, [2350.0]])
    run = runner.RefRun(host)
    from pycuda import driver  # the stand-in (oracle/refshim/pycuda)
    run.setup()
    if workload == "C3":
        return _reference_mixing(run, driver, local)
    run.prepare_flux_solve()
    q = run.q
    npass = (3 if q.singlewalk == 0 else 1000) * int(q.scat) + 1
    points = int(q.nlayer) * int(q.nbin) * int(q.ny) * npass
    flush_buf = driver.mem_alloc(256 << 20)

    def flush():  # 2x L2 overwritten; the reference's 6-7 ms step does not notice the write-back of the dirty lines
        driver.memset_d8(flush_buf, 0, 256 << 20)
        driver.Context.synchronize()

    sampler = ClockSampler(local)
    for _ in range(max(args.warmup, 3)):
        run.flux_solve()
    t = 0.0
    n0 = driver.launches
    for _ in range(args.steps):
        flush()
        t0 = time.perf_counter()  # the reference syncs the device after every launch, so wall clock == its step time
        run.flux_solve()
        t += time.perf_counter() - t0
    launches = driver.launches - n0
    # where the time goes: CUDA events around every launch of two more steps
    driver.PROFILE = True
    driver.per_kernel_ms.clear()
    for _ in range(2):
        flush()
        run.flux_solve()
    driver.PROFILE = False
    split = {k: {"launches_per_step": c // 2, "ms_per_step": ms / 2} for k, (c, ms) in driver.per_kernel_ms.items()}
    clocks = sampler.stop()
    value = points * args.steps / t
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": 1, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": t / args.steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "%s: %d layers x %d bins x %d gauss points, %d fband launches + integrate_flux_double through "
                               "the reference's own computation.py (populate_spectral_flux_iteratively + integrate_flux), "
                               "reference kernels.cu on one B200 (the reference is single-GPU, no CPU path)" %
                               (workload, q.nlayer, q.nbin, q.ny, npass),
                   "l2": "flushed between timed steps (256 MiB device memset)"},
        "kernel_split": split,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": 0, "kind": "reference",
                         "sample": "reference computation.py + kernels.cu (verbatim cubin) on GPU 0; %d launches per step"
                                   % (launches // args.steps)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": int(launches), "clocks": clocks}
    if not args.no_rce and workload in ("C1", "C2"):
        # converged atmospheres per hour through the reference's own radiation_loop + convection_loop (C:827-1174)
        del run
        gc.collect()
        rce = runner.RefRun(_ref_inputs(workload))
        rce.setup()
        res = rce.rce()
        line["rce"] = {"atmospheres_per_hour": 3600.0 / res["seconds"], "seconds": res["seconds"], "status": res["status"],
                       "radiation_iterations": res["radiation_iterations"],
                       "convection_iterations": res["convection_iterations"], "launches": res["launches"],
                       "what": "one %s atmosphere from the isothermal start to the reference's convergence criterion through "
                               "the reference's own radiation_loop + convection_loop, wall clock" % workload}
    emit(line)


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
    ap.add_argument("--workload", default="C2", choices=["C1", "C2", "C3", "C4", "C5"])
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
