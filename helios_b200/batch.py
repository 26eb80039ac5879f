"""Batched grids of independent atmospheres (BASELINE.json configs[4]; SURVEY.md 8e "by atmosphere").

The reference runs one atmosphere per process (`helios.py` is one-atmosphere by design), so a grid of
1024 atmospheres is 1024 runs of ~8 tiny launches per iteration.  Here nbatch atmospheres of identical shape
share ONE launch per kernel: `BatchStore` stacks the per-atmosphere arrays of the individual `Store`s
([b][...], each block with the size the reference allocates), keeps tables and grids once (atmospheres that
use different opacity tables select theirs through `table_index`), and `BatchCompute` drives the same launch
sites as `Compute` with the context in batch mode (helios_ctx_set_batch).  Every atmosphere follows exactly
the trajectory of its own single-atmosphere run: the kernels are deterministic, the iteration counter is
common, and an atmosphere whose nlayer+1 convergence flags are all set is frozen on the device (its
temperature step and flux sweeps are skipped) while the others iterate on.

Supported physics: everything the radiation loop of a premixed run uses (iso / non-iso layers, scattering,
clouds, direct beam without the geometric zenith correction).  Convective adjustment, on-the-fly mixing and
the matrix solver stay single-atmosphere.
"""
import ctypes

import numpy as np

from . import backend
from . import host as _host
from .computation import Compute
from .quantities import Store, _INPUTS, _INPUTS_NONISO, _ZEROS, _DEVICE_ONLY

# host inputs that differ from atmosphere to atmosphere (stacked); everything else in _INPUTS is shared
_PER_ATMOSPHERE = ["p_lay", "p_int", "delta_colmass", "delta_col_upper", "delta_col_lower", "T_lay", "c_p_lay",
                   "kappa_lay", "abs_cross_all_clouds_lay", "scat_cross_all_clouds_lay", "g_0_all_clouds_lay",
                   "abs_cross_all_clouds_int", "scat_cross_all_clouds_int", "g_0_all_clouds_int", "kappa_int",
                   "starflux"]
_TABLES = ["opac_k", "opac_scat_cross", "opac_meanmass"]
# scalars that must agree across the batch (they are passed once per launch)
_COMMON = ["nlayer", "nbin", "ny", "iso", "scat", "clouds", "dir_beam", "scat_corr", "singlewalk", "ntemp", "npress",
           "plancktable_dim", "plancktable_step", "epsi", "epsi2", "mu_star", "g_0", "f_factor", "R_star", "a",
           "R_planet", "T_intern", "real_star", "geom_zenith_corr", "i2s_transition", "adapt_interval", "smooth",
           "foreplay", "physical_tstep", "rad_convergence_limit", "opacity_mixing", "flux_calc_method",
           "energy_correction", "no_atmo_mode", "w_0_limit", "w_0_scat_limit", "delta_tau_limit", "debug", "planet_type"]


class BatchStore(Store):
    """a Store whose per-atmosphere arrays hold nbatch atmospheres"""

    def __init__(self, stores, ctx=None):
        super().__init__(ctx)
        if not stores:
            raise ValueError("empty batch")
        first = stores[0]
        for k, v in vars(first).items():
            if not k.startswith("dev_") and not k.startswith("_"):
                setattr(self, k, v)
        self._ctx = ctx
        self.nbatch = len(stores)
        for name in _COMMON:
            vals = [getattr(s, name) for s in stores]
            if any(v != vals[0] for v in vals[1:]):
                raise ValueError("all atmospheres of a batch must share `%s` (got %r)" % (name, sorted(set(map(str, vals)))))
        if first.opacity_mixing != "premixed" or first.flux_calc_method != "iteration":
            raise ValueError("batched runs support premixed opacities and the iterative flux solver")
        if first.geom_zenith_corr == 1:
            raise ValueError("batched runs do not support the geometric zenith correction")
        if np.dtype(first.fl_prec) != np.float64:
            raise ValueError("helios_b200 implements `precision = double` only")
        for name in _PER_ATMOSPHERE:
            parts = [np.asarray(getattr(s, name), np.float64).reshape(-1) for s in stores if getattr(s, name) is not None]
            if parts:
                if len({p.size for p in parts}) != 1:
                    raise ValueError("`%s` has different sizes across the batch" % name)
                setattr(self, name, np.concatenate(parts))
        # opacity tables: stacked once per distinct table object, selected per atmosphere
        uniq, index = [], []
        for s in stores:
            key = tuple(id(getattr(s, t)) for t in _TABLES)
            hit = [k for k, (kk, _) in enumerate(uniq) if kk == key]
            if hit:
                index.append(hit[0])
            else:
                index.append(len(uniq))
                uniq.append((key, s))
        self.table_index = np.asarray(index, np.int32)
        self.ntables = len(uniq)
        for t in _TABLES:
            setattr(self, t, np.concatenate([np.asarray(getattr(s, t), np.float64).reshape(-1) for _, s in uniq]))
        self.g_batch = np.asarray([float(s.g) for s in stores], np.float64)
        self.T_star_batch = np.asarray([float(s.T_star) for s in stores], np.float64)
        self.F_intern = first.F_intern
        self.dimensions()

    # sizes: every zero / device-only array is per atmosphere, except the shared Planck table
    def _size(self, key):
        n = super()._size(key)
        return n if key == "nplanck_grid" else n * self.nbatch

    def create_zero_arrays(self):
        super().create_zero_arrays()
        self.conv_layer = np.zeros(self.nbatch * (int(self.nlayer) + 1), np.int32)

    def allocate_on_device(self):
        super().allocate_on_device()
        ctx = self.ctx
        self.dev_table_index = ctx.to_device(self.table_index)
        self.dev_g_batch = ctx.to_device(self.g_batch)
        self.dev_planck_star = ctx.zeros(self.nbatch * int(self.nbin))
        self.dev_abort_sums = ctx.zeros(self.nbatch, np.int32)

    def enter(self):
        """put the context into batch mode for this store"""
        ntp = int(self.ntemp) * int(self.npress)
        nb, ny = int(self.nbin), int(self.ny)
        backend._check(backend.lib().helios_ctx_set_batch(
            self.ctx.handle, self.nbatch, int(self.nlayer), nb, ny, ctypes.c_void_p(self.dev_table_index.ptr),
            ntp * nb * ny, ntp * nb, ntp, ctypes.c_void_p(self.dev_g_batch.ptr), ctypes.c_void_p(self.dev_planck_star.ptr)),
            "helios_ctx_set_batch")
        self._entered = True

    def leave(self):
        backend._check(backend.lib().helios_ctx_set_batch(self.ctx.handle, 0, 0, 0, 0, None, 0, 0, 0, None, None),
                       "helios_ctx_set_batch")
        self._entered = False

    def state(self, reset=False):
        """(done flags, iteration counts at which they latched, device iteration counter)"""
        done = np.zeros(self.nbatch, np.int32)
        at = np.zeros(self.nbatch, np.int32)
        it = np.zeros(1, np.int32)
        backend._check(backend.lib().helios_ctx_batch_state(
            self.ctx.handle, done.ctypes.data_as(ctypes.c_void_p), at.ctypes.data_as(ctypes.c_void_p),
            it.ctypes.data_as(ctypes.c_void_p), 1 if reset else 0), "helios_ctx_batch_state")
        return done, at, int(it[0])

    def atmosphere(self, name, b):
        """host copy of atmosphere b's block of a per-atmosphere device array"""
        arr = getattr(self, "dev_" + name).get()
        n = arr.size // self.nbatch
        return arr[b * n:(b + 1) * n]


class BatchCompute(Compute):
    """`Compute`'s launch sites over a BatchStore; the loops are the batched forms of C:827-990"""

    def construct_planck_table(self, quant):
        """shared rows T = 1 + t*step once; one stellar row per atmosphere (K:362-416 with dim = 0 computes
        exactly the row T = T_star)"""
        q = quant
        was_entered = getattr(q, "_entered", False)
        q.leave()  # the set-up kernels are single-atmosphere launches
        self.ctx.call("plancktable", q.dev_planckband_grid, q.dev_opac_interwave, q.dev_opac_deltawave, q.nbin,
                      float(q.T_star_batch[0]), q.plancktable_dim, q.plancktable_step)
        nb = int(q.nbin)
        for b in range(q.nbatch):
            self.ctx.call("plancktable", q.dev_planck_star.view(b * nb, nb), q.dev_opac_interwave, q.dev_opac_deltawave,
                          q.nbin, float(q.T_star_batch[b]), 0, q.plancktable_step)
        if was_entered:
            q.enter()

    def correct_incident_energy(self, quant):
        q = quant
        nb = int(q.nbin)
        was_entered = getattr(q, "_entered", False)
        q.leave()
        for b in range(q.nbatch):
            if q.energy_correction == 1 and q.T_star_batch[b] > 10:
                star = q.dev_starflux.view(b * nb, nb) if q.real_star == 1 else None
                self.ctx.call("corr_inc_energy", q.dev_planck_star.view(b * nb, nb), star, q.dev_opac_deltawave,
                              q.real_star, q.nbin, float(q.T_star_batch[b]), 0, None)
        if was_entered:
            q.enter()

    def _heights(self, quant):
        """H:673-698 per atmosphere (host side, every 10th iteration)"""
        q = quant
        n = int(q.nlayer)
        dz = q.dev_delta_z_lay.get().reshape(q.nbatch, n)
        p = np.asarray(q.p_lay).reshape(q.nbatch, n)
        z = np.zeros((q.nbatch, n))
        one = _Heights()
        one.nlayer, one.planet_type = q.nlayer, q.planet_type
        for b in range(q.nbatch):
            one.delta_z_lay, one.p_lay, one.z_lay = dz[b], p[b], z[b]
            self.hsfunc.calculate_height_z(one)
        q.delta_z_lay = dz.reshape(-1)
        q.z_lay = z.reshape(-1)
        q.dev_z_lay.set(q.z_lay)

    def _refresh_atmosphere(self, quant):
        self.interpolate_opacities_and_scattering_cross_sections(quant)
        self.interpolate_meanmolmass(quant)
        if quant.clouds == 1:
            self.calc_total_g_0_of_gas_and_clouds(quant)
        self.calculate_transmission(quant)
        self.calculate_delta_z(quant)
        self._heights(quant)
        self.calculate_direct_beamflux(quant)
        self.build_flux_plan(quant)

    def _iteration(self, quant, refresh, heights=True, fused=True):
        """one RT iteration of C:851-984 on the device (the temperature step reads the device iteration counter).
        fused: temp_inter + both Planck interpolations are one launch, and temperature step + flag sum + latch +
        counter advance are one launch (5 launches per iteration instead of 9); bitwise the same results."""
        q = quant
        if fused:
            self.prepare_iteration(q)
        else:
            self.interpolate_temperatures(q)
            self.interpolate_planck(q)
        if refresh:
            self.interpolate_opacities_and_scattering_cross_sections(q)
            self.interpolate_meanmolmass(q)
            if q.clouds == 1:
                self.calc_total_g_0_of_gas_and_clouds(q)
            self.calculate_transmission(q)
            self.calculate_delta_z(q)
            if heights:
                self._heights(q)
            self.calculate_direct_beamflux(q)
            self.build_flux_plan(q)
        self.populate_spectral_flux_iteratively(q)
        self.integrate_flux(q)
        if fused:
            self.ctx.call("rad_temp_iter_latched", q.dev_F_down_tot, q.dev_F_up_tot, q.dev_F_net, q.dev_F_net_diff,
                          q.dev_T_lay, q.dev_p_lay, q.dev_T_int, q.dev_p_int, q.dev_abort, q.dev_T_store,
                          q.dev_delta_t_prefactor, q.dev_F_add_heat_lay, q.dev_F_add_heat_sum, q.dev_F_smooth,
                          q.dev_F_smooth_sum, q.dev_c_p_lay, q.dev_meanmolmass_lay, q.iter_value, q.f_factor,
                          q.foreplay, q.g, q.nlayer, q.physical_tstep, q.rad_convergence_limit, q.adapt_interval,
                          q.smooth, q.plancktable_dim, q.plancktable_step, q.F_intern, q.no_atmo_mode,
                          q.dev_abort_sums)
        else:
            self.rad_temp_iteration(q)
            self.ctx.call("abort_sum", q.dev_abort, int(q.nlayer) + 1, q.dev_abort_sums)
            self.ctx.call("batch_iter_advance")

    def radiation_loop(self, quant, write=None, read=None, rt_plot=None, graph=True, fused=True):
        """C:851-984 for nbatch atmospheres at once, bookkeeping on the device.

        The iteration counter, the per-atmosphere convergence latch and the iteration count at which each
        atmosphere converged live on the device, so a block of 10 iterations (one opacity refresh + 10 flux
        solves and temperature steps, the reference's schedule C:860) is replayed as ONE CUDA graph and the host
        looks at the state once per block.  Converged atmospheres are frozen by the latch, so the (up to 9)
        iterations a block runs past an atmosphere's convergence do not touch it: `converged_at[b]` and the final
        state equal `iter_value` and the state of its single-atmosphere run.  The altitude grid (host side,
        H:673) feeds no kernel without the geometric zenith correction and is evaluated once at the end."""
        q = quant
        if int(q.foreplay) != 0 or q.physical_tstep != 0 or q.singlewalk == 1:
            raise ValueError("the batched loop covers the default iterative run (foreplay = 0, no physical timestep)")
        # The host looks at the state once per block of 10 iterations, so the criterion relaxation of C:974 (applied
        # when iter_value == n exactly) can only be reproduced for relaxation numbers on a block boundary; the default
        # param.dat uses 1e4 and 2e4.  Anything else would switch up to 9 iterations late -> refuse instead of drifting.
        # Not reproduced here (documented deviation): the reference's escape to the convection loop when the surface
        # temperature leaves the Planck table (C:946-952) -- the temperature step clamps to the table instead.
        bad = [n for n in (q.crit_relaxation_numbers or []) if int(n) % 10 != 0]
        if bad:
            raise ValueError("the batched loop needs crit_relaxation_numbers that are multiples of 10, got %r" % bad)
        q.enter()
        q.iter_value = np.int32(0)  # the kernels read the device counter; the argument is ignored
        lib = backend.lib()
        backend._check(lib.helios_ctx_batch_device_iteration(self.ctx.handle, 1), "helios_ctx_batch_device_iteration")
        q.state(reset=True)
        ev = (self.ctx.event(), self.ctx.event())
        ev[0].record()
        block = 10
        graphs = {}
        import time as _time
        wall = {"first_block_s": 0.0, "capture_s": 0.0, "replay_s": 0.0}
        t_mark = _time.perf_counter()
        try:
            # the first block runs eagerly: it sizes the library's scratch buffers, which must not grow while capturing
            for k in range(block):
                self._iteration(q, refresh=(k == 0), heights=False, fused=fused)
            done, at, it = q.state()
            wall["first_block_s"] = _time.perf_counter() - t_mark
            t_mark = _time.perf_counter()
            while not done.all():
                limit = float(q.rad_convergence_limit)
                if graph:
                    if limit not in graphs:
                        t_cap = _time.perf_counter()
                        with self.ctx.capture() as g:
                            for k in range(block):
                                self._iteration(q, refresh=(k == 0), heights=False, fused=fused)
                        graphs[limit] = g
                        wall["capture_s"] += _time.perf_counter() - t_cap
                    graphs[limit].launch()
                else:
                    for k in range(block):
                        self._iteration(q, refresh=(k == 0), heights=False, fused=fused)
                done, at, it = q.state()
                if self.verbose and it % 100 == 0:
                    print("batch iteration %d: %d of %d atmospheres converged" % (it, int(done.sum()), q.nbatch))
                for n_relax in (q.crit_relaxation_numbers or []):
                    if it - block < n_relax <= it:
                        self.hsfunc.relax_radiative_convergence_criterion(q)
                if it > q.max_nr_iterations:
                    print("\nRun exceeds allowed maximum allowed number of iteration steps. Aborting...")
                    raise SystemExit()
            wall["replay_s"] = _time.perf_counter() - t_mark - wall["capture_s"]
            q.converged_at = at.astype(np.int64)
            q.iter_value = np.int32(it)
            self._heights(q)
            self.stats["radiation_loop_wall"] = wall
        finally:
            ev[1].record()
            ev[1].synchronize()
            self.stats["radiation_loop_ms"] = ev[0].time_till(ev[1])
            self.stats["radiation_iterations"] = int(q.iter_value)
            backend._check(lib.helios_ctx_batch_device_iteration(self.ctx.handle, 0), "helios_ctx_batch_device_iteration")
            q.leave()

    def convection_loop(self, quant, write=None, read=None, rt_plot=None):
        raise NotImplementedError("convective adjustment is single-atmosphere (host_functions.py works on one profile)")


class _Heights(object):
    pass


def make_batch(stores, ctx):
    """stack, upload, and return (BatchStore, BatchCompute)"""
    q = BatchStore(stores, ctx)
    q.create_zero_arrays()
    q.convert_input_list_to_array()
    q.copy_host_to_device()
    q.allocate_on_device()
    return q, BatchCompute(ctx, verbose=False)
