"""Oracle A: the reference's own kernels.cu, compiled verbatim (oracle/Makefile -> oracle/_ref/helios_ref.cubin)
and launched through the CUDA driver API with the exact block/grid shapes of source/computation.py.
TEST INFRASTRUCTURE, NOT PRODUCT -- used only by tests/ and by `bench.py --impl reference`.

This is PyCUDA's `SourceModule.get_function(name)(*args, block=, grid=)` without PyCUDA (which is not
installed here): np.int32 -> int, np.float64 -> double, device arrays -> pointers, then
cuLaunchKernel on the NULL stream followed by cuCtxSynchronize, as computation.py does after every
launch (`cuda.Context.synchronize()`, C:60 ...).  Device memory is the product library's: the cubin
is loaded into the device's primary context, the same one the CUDA runtime inside libhelios_b200.so
uses, so pointers are interchangeable.
"""
import ctypes
import os

import numpy as np

CUBIN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "helios_ref.cubin")


def available():
    return os.path.exists(CUBIN)


class RefError(RuntimeError):
    pass


class RefModule:
    def __init__(self, device=0, cubin=CUBIN):
        if not os.path.exists(cubin):
            raise RefError("reference cubin %s missing: run `make -C oracle ref` where /root/reference exists" % cubin)
        self.cu = ctypes.CDLL("libcuda.so.1")
        self._chk(self.cu.cuInit(0), "cuInit")
        dev = ctypes.c_int()
        self._chk(self.cu.cuDeviceGet(ctypes.byref(dev), int(device)), "cuDeviceGet")
        self.ctx = ctypes.c_void_p()
        self._chk(self.cu.cuDevicePrimaryCtxRetain(ctypes.byref(self.ctx), dev), "cuDevicePrimaryCtxRetain")
        self._chk(self.cu.cuCtxSetCurrent(self.ctx), "cuCtxSetCurrent")
        image = open(cubin, "rb").read()
        self._image = ctypes.create_string_buffer(image, len(image))
        self.mod = ctypes.c_void_p()
        self._chk(self.cu.cuModuleLoadData(ctypes.byref(self.mod), self._image), "cuModuleLoadData")
        self._fn = {}
        self.launches = 0

    def _chk(self, rc, what):
        if rc != 0:
            name = ctypes.c_char_p()
            try:
                self.cu.cuGetErrorName(rc, ctypes.byref(name))
            except Exception:
                pass
            raise RefError("%s failed: CUresult %d (%s)" % (what, rc, name.value.decode() if name.value else "?"))

    def get_function(self, name):
        if name not in self._fn:
            f = ctypes.c_void_p()
            self._chk(self.cu.cuModuleGetFunction(ctypes.byref(f), self.mod, name.encode()), "cuModuleGetFunction(%s)" % name)
            self._fn[name] = f
        fn = self._fn[name]

        def launch(*args, block, grid, sync=True):
            holders = []
            for a in args:
                if a is None:
                    holders.append(ctypes.c_void_p(0))
                elif isinstance(a, np.integer):
                    holders.append(ctypes.c_int(int(a)))
                elif isinstance(a, (float, np.floating)):
                    holders.append(ctypes.c_double(float(a)))
                elif hasattr(a, "ptr"):
                    holders.append(ctypes.c_void_p(int(a.ptr)))
                elif isinstance(a, int):
                    # plain Python ints are device addresses (PyCUDA takes scalars as numpy types only)
                    holders.append(ctypes.c_void_p(a))
                else:
                    raise TypeError("unsupported kernel argument %r" % (a,))
            params = (ctypes.c_void_p * len(holders))(*[ctypes.cast(ctypes.byref(h), ctypes.c_void_p) for h in holders])
            self._chk(self.cu.cuCtxSetCurrent(self.ctx), "cuCtxSetCurrent")
            g = tuple(int(v) for v in grid) + (1,) * (3 - len(grid))
            b = tuple(int(v) for v in block) + (1,) * (3 - len(block))
            self._chk(self.cu.cuLaunchKernel(fn, g[0], g[1], g[2], b[0], b[1], b[2], 0, None, params, None),
                      "cuLaunchKernel(%s)" % name)
            self.launches += 1
            if sync:
                self._chk(self.cu.cuCtxSynchronize(), "cuCtxSynchronize after %s" % name)

        return launch

    def synchronize(self):
        self._chk(self.cu.cuCtxSynchronize(), "cuCtxSynchronize")


def i32(v):
    return np.int32(v)


def f64(v):
    return np.float64(v)


class RefCompute:
    """The launch sites of source/computation.py (block/grid shapes and argument order), re-pointed
    at RefModule.  `q` is any object carrying the `dev_*` buffers and scalars of quantities.Store."""

    def __init__(self, device=0):
        self.mod = RefModule(device)

    def _k(self, name):
        return self.mod.get_function(name)

    def construct_planck_table(self, q):  # C:39-60
        for p_iter in range(10):
            self._k("plancktable")(q.dev_planckband_grid, q.dev_opac_interwave, q.dev_opac_deltawave, i32(q.nbin),
                                   f64(q.T_star), i32(p_iter), i32(q.plancktable_dim), i32(q.plancktable_step),
                                   block=(16, 16, 1),
                                   grid=((int(q.nbin) + 15) // 16, (int(q.plancktable_dim / 10 + 1) + 15) // 16, 1))

    def correct_incident_energy(self, q):  # C:62-82
        if q.energy_correction == 1 and q.T_star > 10:
            self._k("corr_inc_energy")(q.dev_planckband_grid, q.dev_starflux, q.dev_opac_deltawave, i32(q.real_star),
                                       i32(q.nbin), f64(q.T_star), i32(q.plancktable_dim), block=(16, 1, 1),
                                       grid=((int(q.nbin) + 15) // 16, 1, 1))

    def interpolate_temperatures(self, q):  # C:104-117
        self._k("temp_inter")(q.dev_T_lay, q.dev_T_int, i32(q.ninterface), i32(q.iter_value), block=(16, 1, 1),
                              grid=((int(q.ninterface) + 15) // 16, 1, 1))

    def interpolate_planck(self, q):  # C:294-329
        self._k("planck_interpol_layer")(q.dev_T_lay, q.dev_planckband_lay, q.dev_planckband_grid, q.dev_starflux,
                                         i32(q.real_star), i32(q.nlayer), i32(q.nbin), i32(q.plancktable_dim),
                                         i32(q.plancktable_step), block=(16, 16, 1),
                                         grid=((int(q.nbin) + 15) // 16, (int(q.nlayer + 2) + 15) // 16, 1))
        if q.iso == 0:
            self._k("planck_interpol_interface")(q.dev_T_int, q.dev_planckband_int, q.dev_planckband_grid,
                                                 i32(q.ninterface), i32(q.nbin), i32(q.plancktable_dim),
                                                 i32(q.plancktable_step), block=(16, 16, 1),
                                                 grid=((int(q.nbin) + 15) // 16, (int(q.ninterface) + 15) // 16, 1))

    def interpolate_opacities_and_scattering_cross_sections(self, q):  # C:119-161
        k = self._k("opac_interpol")
        k(q.dev_T_lay, q.dev_ktemp, q.dev_p_lay, q.dev_kpress, q.dev_opac_k, q.dev_opac_wg_lay, q.dev_opac_scat_cross,
          q.dev_scat_cross_lay, i32(q.npress), i32(q.ntemp), i32(q.ny), i32(q.nbin), i32(q.nlayer), block=(16, 16, 1),
          grid=((int(q.nbin) + 15) // 16, (int(q.nlayer) + 15) // 16, 1))
        if q.iso == 0:
            k(q.dev_T_int, q.dev_ktemp, q.dev_p_int, q.dev_kpress, q.dev_opac_k, q.dev_opac_wg_int,
              q.dev_opac_scat_cross, q.dev_scat_cross_int, i32(q.npress), i32(q.ntemp), i32(q.ny), i32(q.nbin),
              i32(q.ninterface), block=(16, 16, 1), grid=((int(q.nbin) + 15) // 16, (int(q.ninterface) + 15) // 16, 1))

    def interpolate_meanmolmass(self, q):  # C:163-197
        k = self._k("meanmolmass_interpol")
        k(q.dev_T_lay, q.dev_ktemp, q.dev_meanmolmass_lay, q.dev_opac_meanmass, q.dev_p_lay, q.dev_kpress,
          i32(q.npress), i32(q.ntemp), i32(q.nlayer), block=(16, 1, 1), grid=((int(q.nlayer) + 15) // 16, 1, 1))
        if q.iso == 0:
            k(q.dev_T_int, q.dev_ktemp, q.dev_meanmolmass_int, q.dev_opac_meanmass, q.dev_p_int, q.dev_kpress,
              i32(q.npress), i32(q.ntemp), i32(q.ninterface), block=(16, 1, 1),
              grid=((int(q.ninterface) + 15) // 16, 1, 1))

    def calc_total_g_0_of_gas_and_clouds(self, q):  # C:331-362
        k = self._k("calc_total_g_0_of_gas_and_clouds")
        k(q.dev_scat_cross_lay, q.dev_g_0_all_clouds_lay, q.dev_scat_cross_all_clouds_lay, q.dev_g_0_tot_lay,
          f64(q.g_0), i32(q.nbin), i32(q.nlayer), block=(16, 16, 1),
          grid=((int(q.nbin) + 15) // 16, (int(q.nlayer) + 15) // 16, 1))
        if q.iso == 0:
            k(q.dev_scat_cross_int, q.dev_g_0_all_clouds_int, q.dev_scat_cross_all_clouds_int, q.dev_g_0_tot_int,
              f64(q.g_0), i32(q.nbin), i32(q.ninterface), block=(16, 16, 1),
              grid=((int(q.nbin) + 15) // 16, (int(q.ninterface) + 15) // 16, 1))

    def calculate_transmission(self, q):  # C:364-462
        q.dev_scat_trigger.fill_zero()
        q.dev_scat_trigger.ctx.synchronize()
        grid = ((int(q.nbin) + 15) // 16, (int(q.ny) + 3) // 4, (int(q.nlayer) + 3) // 4)
        tail = (f64(q.g_0), f64(q.epsi), f64(q.epsi2), f64(q.mu_star), f64(q.w_0_limit), f64(q.w_0_scat_limit),
                i32(q.scat), i32(q.nbin), i32(q.ny), i32(q.nlayer), i32(q.clouds), i32(q.scat_corr), i32(q.debug),
                f64(q.i2s_transition))
        if q.iso == 1:
            self._k("calc_trans_iso")(q.dev_trans_wg, q.dev_delta_tau_wg, q.dev_M_term, q.dev_N_term, q.dev_P_term,
                                      q.dev_G_plus, q.dev_G_minus, q.dev_delta_colmass, q.dev_opac_wg_lay,
                                      q.dev_meanmolmass_lay, q.dev_scat_cross_lay, q.dev_abs_cross_all_clouds_lay,
                                      q.dev_scat_cross_all_clouds_lay, q.dev_delta_tau_all_clouds, q.dev_w_0,
                                      q.dev_g_0_tot_lay, q.dev_scat_trigger, *tail, block=(16, 4, 4), grid=grid)
        else:
            self._k("calc_trans_noniso")(
                q.dev_trans_wg_upper, q.dev_trans_wg_lower, q.dev_delta_tau_wg_upper, q.dev_delta_tau_wg_lower,
                q.dev_M_upper, q.dev_M_lower, q.dev_N_upper, q.dev_N_lower, q.dev_P_upper, q.dev_P_lower,
                q.dev_G_plus_upper, q.dev_G_plus_lower, q.dev_G_minus_upper, q.dev_G_minus_lower,
                q.dev_delta_col_upper, q.dev_delta_col_lower, q.dev_opac_wg_lay, q.dev_opac_wg_int,
                q.dev_meanmolmass_lay, q.dev_meanmolmass_int, q.dev_scat_cross_lay, q.dev_scat_cross_int,
                q.dev_abs_cross_all_clouds_lay, q.dev_abs_cross_all_clouds_int, q.dev_scat_cross_all_clouds_lay,
                q.dev_scat_cross_all_clouds_int, q.dev_delta_tau_all_clouds_upper, q.dev_delta_tau_all_clouds_lower,
                q.dev_w_0_upper, q.dev_w_0_lower, q.dev_g_0_tot_lay, q.dev_g_0_tot_int, q.dev_scat_trigger, *tail,
                block=(16, 4, 4), grid=grid)

    def calculate_delta_z(self, q):  # C:464-479
        self._k("calc_delta_z")(q.dev_T_lay, q.dev_p_int, q.dev_p_lay, q.dev_meanmolmass_lay, q.dev_delta_z_lay,
                                f64(q.g), i32(q.nlayer), block=(16, 1, 1), grid=((int(q.nlayer) + 15) // 16, 1, 1))

    def calculate_direct_beamflux(self, q):  # C:481-526
        grid = ((int(q.ninterface) + 3) // 4, (int(q.nbin) + 31) // 32, (int(q.ny) + 3) // 4)
        tail = (q.dev_z_lay, f64(q.mu_star), f64(q.R_planet), f64(q.R_star), f64(q.a), i32(q.dir_beam),
                i32(q.geom_zenith_corr), i32(q.ninterface), i32(q.nbin), i32(q.ny))
        if q.iso == 1:
            self._k("fdir_iso")(q.dev_F_dir_wg, q.dev_planckband_lay, q.dev_delta_tau_wg, *tail, block=(4, 32, 4), grid=grid)
        else:
            self._k("fdir_noniso")(q.dev_F_dir_wg, q.dev_Fc_dir_wg, q.dev_planckband_lay, q.dev_delta_tau_wg_upper,
                                   q.dev_delta_tau_wg_lower, *tail, block=(4, 32, 4), grid=grid)

    def populate_spectral_flux_iteratively(self, q, sync_each=True):  # C:528-623
        nscat_step = 3 if q.singlewalk == 0 else 1000
        grid = ((int(q.nbin) + 15) // 16, (int(q.ny) + 15) // 16, 1)
        for _ in range(nscat_step * int(q.scat) + 1):
            if q.iso == 1:
                self._k("fband_iso")(
                    q.dev_F_down_wg, q.dev_F_up_wg, q.dev_F_dir_wg, q.dev_planckband_lay, q.dev_w_0, q.dev_M_term,
                    q.dev_N_term, q.dev_P_term, q.dev_G_plus, q.dev_G_minus, q.dev_surf_albedo, q.dev_g_0_tot_lay,
                    f64(q.g_0), i32(q.singlewalk), f64(q.R_star), f64(q.a), i32(q.ninterface), i32(q.nbin),
                    f64(q.f_factor), f64(q.mu_star), i32(q.ny), f64(q.epsi), i32(q.dir_beam), i32(q.clouds),
                    i32(q.scat_corr), i32(q.debug), f64(q.i2s_transition), block=(16, 16, 1), grid=grid, sync=sync_each)
            else:
                self._k("fband_noniso")(
                    q.dev_F_down_wg, q.dev_F_up_wg, q.dev_Fc_down_wg, q.dev_Fc_up_wg, q.dev_F_dir_wg, q.dev_Fc_dir_wg,
                    q.dev_planckband_lay, q.dev_planckband_int, q.dev_w_0_upper, q.dev_w_0_lower,
                    q.dev_delta_tau_wg_upper, q.dev_delta_tau_wg_lower, q.dev_delta_tau_all_clouds_upper,
                    q.dev_delta_tau_all_clouds_lower, q.dev_M_upper, q.dev_M_lower, q.dev_N_upper, q.dev_N_lower,
                    q.dev_P_upper, q.dev_P_lower, q.dev_G_plus_upper, q.dev_G_plus_lower, q.dev_G_minus_upper,
                    q.dev_G_minus_lower, q.dev_surf_albedo, q.dev_g_0_tot_lay, q.dev_g_0_tot_int, f64(q.g_0),
                    i32(q.singlewalk), f64(q.R_star), f64(q.a), i32(q.ninterface), i32(q.nbin), f64(q.f_factor),
                    f64(q.mu_star), i32(q.ny), f64(q.epsi), f64(q.delta_tau_limit), i32(q.dir_beam), i32(q.clouds),
                    i32(q.scat_corr), i32(q.debug), f64(q.i2s_transition), block=(16, 16, 1), grid=grid, sync=sync_each)
        if not sync_each:
            self.mod.synchronize()

    def solve_for_spectral_fluxes_via_matrix(self, q):  # C:625-729
        grid = ((int(q.nbin) + 15) // 16, (int(q.ny) + 15) // 16, 1)
        if q.iso == 1:
            self._k("fband_matrix_iso")(
                q.dev_F_down_wg, q.dev_F_up_wg, q.dev_F_dir_wg, q.dev_planckband_lay, q.dev_w_0, q.dev_M_term,
                q.dev_N_term, q.dev_P_term, q.dev_G_plus, q.dev_G_minus, q.dev_g_0_tot_lay, q.dev_alpha, q.dev_beta,
                q.dev_source_term_down, q.dev_source_term_up, q.dev_c_prime, q.dev_d_prime, q.dev_scat_trigger,
                q.dev_trans_wg, q.dev_surf_albedo, f64(q.g_0), i32(q.singlewalk), f64(q.R_star), f64(q.a),
                i32(q.ninterface), i32(q.nbin), f64(q.f_factor), f64(q.mu_star), i32(q.ny), f64(q.epsi),
                i32(q.dir_beam), i32(q.clouds), i32(q.scat_corr), i32(q.debug), f64(q.i2s_transition),
                block=(16, 16, 1), grid=grid)
        else:
            self._k("fband_matrix_noniso")(
                q.dev_F_down_wg, q.dev_F_up_wg, q.dev_Fc_down_wg, q.dev_Fc_up_wg, q.dev_F_dir_wg, q.dev_Fc_dir_wg,
                q.dev_planckband_lay, q.dev_planckband_int, q.dev_w_0_upper, q.dev_w_0_lower,
                q.dev_delta_tau_wg_upper, q.dev_delta_tau_wg_lower, q.dev_delta_tau_all_clouds_upper,
                q.dev_delta_tau_all_clouds_lower, q.dev_M_upper, q.dev_M_lower, q.dev_N_upper, q.dev_N_lower,
                q.dev_P_upper, q.dev_P_lower, q.dev_G_plus_upper, q.dev_G_plus_lower, q.dev_G_minus_upper,
                q.dev_G_minus_lower, q.dev_g_0_tot_lay, q.dev_g_0_tot_int, q.dev_alpha, q.dev_beta,
                q.dev_source_term_down, q.dev_source_term_up, q.dev_c_prime, q.dev_d_prime, q.dev_scat_trigger,
                q.dev_trans_wg_upper, q.dev_trans_wg_lower, q.dev_surf_albedo, f64(q.g_0), i32(q.singlewalk),
                f64(q.R_star), f64(q.a), i32(q.ninterface), i32(q.nbin), f64(q.f_factor), f64(q.mu_star), i32(q.ny),
                f64(q.epsi), f64(q.delta_tau_limit), i32(q.dir_beam), i32(q.clouds), i32(q.scat_corr), i32(q.debug),
                f64(q.i2s_transition), block=(16, 16, 1), grid=grid)

    def integrate_flux(self, q):  # C:731-757
        self._k("integrate_flux_double")(q.dev_opac_deltawave, q.dev_F_down_tot, q.dev_F_up_tot, q.dev_F_net,
                                         q.dev_F_down_wg, q.dev_F_up_wg, q.dev_F_dir_wg, q.dev_F_down_band,
                                         q.dev_F_up_band, q.dev_F_dir_band, q.dev_gauss_weight, i32(q.nbin),
                                         i32(q.ninterface), i32(q.ny), block=(32, 4, 8), grid=(1, 1, 1))

    def rad_temp_iteration(self, q):  # C:759-797
        self._k("rad_temp_iter")(
            q.dev_F_down_tot, q.dev_F_up_tot, q.dev_F_net, q.dev_F_net_diff, q.dev_T_lay, q.dev_p_lay, q.dev_T_int,
            q.dev_p_int, q.dev_abort, q.dev_T_store, q.dev_delta_t_prefactor, q.dev_F_add_heat_lay,
            q.dev_F_add_heat_sum, q.dev_F_smooth, q.dev_F_smooth_sum, q.dev_c_p_lay, q.dev_meanmolmass_lay,
            i32(q.iter_value), f64(q.f_factor), i32(q.foreplay), f64(q.g), i32(q.nlayer), f64(q.physical_tstep),
            f64(q.rad_convergence_limit), i32(q.adapt_interval), i32(q.smooth), i32(q.plancktable_dim),
            i32(q.plancktable_step), f64(q.F_intern), i32(q.no_atmo_mode), block=(16, 1, 1),
            grid=((int(q.nlayer + 1) + 15) // 16, 1, 1))

    def conv_temp_iteration(self, q):  # C:799-825
        self._k("conv_temp_iter")(
            q.dev_F_down_tot, q.dev_F_up_tot, q.dev_F_net, q.dev_F_net_diff, q.dev_T_lay, q.dev_p_lay, q.dev_p_int,
            q.dev_T_store, q.dev_delta_t_prefactor, q.dev_marked_red, q.dev_F_add_heat_lay, q.dev_F_smooth,
            q.dev_F_smooth_sum, i32(q.nlayer), i32(q.iter_value), i32(q.adapt_interval), i32(q.smooth),
            f64(q.F_intern), block=(16, 1, 1), grid=((int(q.nlayer + 1) + 15) // 16, 1, 1))

    def interpolate_species_opac(self, q):  # C:1298-1336
        k = self._k("opac_species_interpol")
        k(q.dev_T_lay, q.dev_ktemp, q.dev_p_lay, q.dev_kpress, q.dev_opacity_spec_pretab, q.dev_opac_spec_wg_lay,
          i32(q.npress), i32(q.ntemp), i32(q.ny), i32(q.nbin), i32(q.nlayer), block=(16, 16, 1),
          grid=((int(q.nbin) + 15) // 16, (int(q.nlayer) + 15) // 16, 1))
        if q.iso == 0:
            k(q.dev_T_int, q.dev_ktemp, q.dev_p_int, q.dev_kpress, q.dev_opacity_spec_pretab, q.dev_opac_spec_wg_int,
              i32(q.npress), i32(q.ntemp), i32(q.ny), i32(q.nbin), i32(q.ninterface), block=(16, 16, 1),
              grid=((int(q.nbin) + 15) // 16, (int(q.ninterface) + 15) // 16, 1))

    def add_to_mixed_opacity(self, q, mass_spec_g, s, ro_method):  # C:1338-1388 (mass already in grams)
        k = self._k("add_to_mixed_opac")
        k(q.dev_vmr_spec_lay, q.dev_opac_spec_wg_lay, q.dev_opac_wg_lay, q.dev_meanmolmass_lay, q.dev_gauss_weight,
          q.dev_gauss_y, f64(mass_spec_g), i32(s), i32(ro_method), i32(q.ny), i32(q.nbin), i32(q.nlayer),
          block=(32, 32, 1), grid=((int(q.nbin) + 31) // 32, (int(q.nlayer) + 31) // 32, 1))
        if q.iso == 0:
            k(q.dev_vmr_spec_int, q.dev_opac_spec_wg_int, q.dev_opac_wg_int, q.dev_meanmolmass_int, q.dev_gauss_weight,
              q.dev_gauss_y, f64(mass_spec_g), i32(s), i32(ro_method), i32(q.ny), i32(q.nbin), i32(q.ninterface),
              block=(32, 32, 1), grid=((int(q.nbin) + 31) // 32, (int(q.ninterface) + 31) // 32, 1))

    def calculate_H2O_Rayleigh_scattering(self, q, s):  # C:1390-1423
        from helios_b200 import host
        mass = f64(q.species_list[s].weight * host.AMU)
        k = self._k("calc_h2o_scat")
        k(q.dev_T_lay, q.dev_p_lay, q.dev_opac_wave, q.dev_scat_cross_spec_lay, q.dev_vmr_spec_lay, mass, i32(q.nbin),
          i32(q.nlayer), block=(16, 16, 1), grid=((int(q.nbin) + 15) // 16, (int(q.nlayer) + 15) // 16, 1))
        if q.iso == 0:
            k(q.dev_T_int, q.dev_p_int, q.dev_opac_wave, q.dev_scat_cross_spec_int, q.dev_vmr_spec_int, mass,
              i32(q.nbin), i32(q.ninterface), block=(16, 16, 1),
              grid=((int(q.nbin) + 15) // 16, (int(q.ninterface) + 15) // 16, 1))

    def add_to_mixed_scat_cross_sect(self, q):  # C:1425-1452
        k = self._k("add_to_mixed_scat")
        k(q.dev_vmr_spec_lay, q.dev_scat_cross_spec_lay, q.dev_scat_cross_lay, i32(q.nbin), i32(q.nlayer),
          block=(16, 16, 1), grid=((int(q.nbin) + 15) // 16, (int(q.nlayer) + 15) // 16, 1))
        if q.iso == 0:
            k(q.dev_vmr_spec_int, q.dev_scat_cross_spec_int, q.dev_scat_cross_int, i32(q.nbin), i32(q.ninterface),
              block=(16, 16, 1), grid=((int(q.nbin) + 15) // 16, (int(q.ninterface) + 15) // 16, 1))

    # ---- kappa / c_p / entropy / phase (C:199-292): launched only when kappa comes from a file
    def _entr(self, name, temp, press, out, tab, n, q):
        self._k(name)(temp, q.dev_entr_temp, press, q.dev_entr_press, out, tab, i32(q.entr_npress), i32(q.entr_ntemp),
                      i32(n), block=(16, 1, 1), grid=((int(n) + 15) // 16, 1, 1))

    def interpolate_kappa_and_cp(self, q):
        if not isinstance(q.input_kappa_value, str):
            return
        self._entr("kappa_interpol", q.dev_T_lay, q.dev_p_lay, q.dev_kappa_lay, q.dev_entr_kappa, q.nlayer, q)
        self._entr("cp_interpol", q.dev_T_lay, q.dev_p_lay, q.dev_c_p_lay, q.dev_entr_c_p, q.nlayer, q)
        if q.iso == 0:
            self._entr("kappa_interpol", q.dev_T_int, q.dev_p_int, q.dev_kappa_int, q.dev_entr_kappa, q.ninterface, q)

    def interpolate_entropy(self, q):
        if isinstance(q.input_kappa_value, str):
            self._entr("entropy_interpol", q.dev_T_lay, q.dev_p_lay, q.dev_entropy_lay, q.dev_entr_entropy, q.nlayer, q)

    def interpolate_phase_state(self, q):
        if q.input_kappa_value == "water_atmo":
            self._entr("phase_number_interpol", q.dev_T_lay, q.dev_p_lay, q.dev_phase_number_lay,
                       q.dev_entr_phase_number, q.nlayer, q)

    # ---- post-processing (C:1176-1296)
    def integrate_optdepth_transmission(self, q):
        grid = ((int(q.nbin) + 15) // 16, (int(q.nlayer) + 15) // 16, 1)
        if q.iso == 1:
            self._k("integrate_optdepth_transmission_iso")(
                q.dev_trans_wg, q.dev_trans_band, q.dev_delta_tau_wg, q.dev_delta_tau_band, q.dev_gauss_weight,
                i32(q.nbin), i32(q.nlayer), i32(q.ny), block=(16, 16, 1), grid=grid)
        else:
            self._k("integrate_optdepth_transmission_noniso")(
                q.dev_trans_wg_upper, q.dev_trans_wg_lower, q.dev_trans_band, q.dev_delta_tau_wg_upper,
                q.dev_delta_tau_wg_lower, q.dev_delta_tau_band, q.dev_gauss_weight, q.dev_delta_tau_all_clouds,
                q.dev_delta_tau_all_clouds_upper, q.dev_delta_tau_all_clouds_lower, i32(q.nbin), i32(q.nlayer),
                i32(q.ny), block=(16, 16, 1), grid=grid)

    def calculate_contribution_function(self, q):
        grid = ((int(q.nbin) + 15) // 16, (int(q.nlayer) + 15) // 16, 1)
        if q.iso == 1:
            self._k("calc_contr_func_iso")(
                q.dev_trans_wg, q.dev_trans_weight_band, q.dev_contr_func_band, q.dev_gauss_weight,
                q.dev_planckband_lay, f64(q.epsi), i32(q.nbin), i32(q.nlayer), i32(q.ny), block=(16, 16, 1), grid=grid)
        else:
            self._k("calc_contr_func_noniso")(
                q.dev_trans_wg_upper, q.dev_trans_wg_lower, q.dev_trans_weight_band, q.dev_contr_func_band,
                q.dev_gauss_weight, q.dev_planckband_lay, f64(q.epsi), i32(q.nbin), i32(q.nlayer), i32(q.ny),
                block=(16, 16, 1), grid=grid)

    def calculate_mean_opacities(self, q):
        self._k("calc_mean_opacities")(
            q.dev_planck_opac_T_pl, q.dev_ross_opac_T_pl, q.dev_planck_opac_T_star, q.dev_ross_opac_T_star,
            q.dev_opac_wg_lay, q.dev_abs_cross_all_clouds_lay, q.dev_meanmolmass_lay, q.dev_planckband_lay,
            q.dev_opac_interwave, q.dev_opac_deltawave, q.dev_T_lay, q.dev_gauss_weight, q.dev_gauss_y,
            q.dev_opac_band_lay, i32(q.nlayer), i32(q.nbin), i32(q.ny), f64(q.T_star), block=(16, 1, 1),
            grid=((int(q.nlayer) + 15) // 16, 1, 1))

    def integrate_beamflux(self, q):
        self._k("integrate_beamflux")(q.dev_F_dir_tot, q.dev_F_dir_band, q.dev_opac_deltawave, q.dev_gauss_weight,
                                      i32(q.nbin), i32(q.ninterface), block=(16, 1, 1),
                                      grid=((int(q.ninterface) + 15) // 16, 1, 1))

    # generic access for the remaining (post-processing) kernels
    def launch(self, name, *args, block, grid):
        self._k(name)(*args, block=block, grid=grid)
