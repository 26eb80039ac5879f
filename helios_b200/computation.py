"""Host-side mirror of the reference's `computation.Compute` (source/computation.py:31-1501).

Same method names, same arguments, same loop semantics -- but every launch site goes through the
C-ABI of libhelios_b200.so instead of `SourceModule.get_function(...)(..., block=, grid=)`, nothing is
JIT-compiled at start-up (C:34-37), launches are stream-ordered instead of followed by a full device
sync, the (3*scat+1) flux passes are one fused launch, and the per-iteration convergence test is a
4-byte read of a device-side sum instead of a D2H of the flag array plus a Python loop (C:927-932).

`hsfunc` is the module with the host-side helpers: the reference's unchanged
`source.host_functions` in drop-in mode, `helios_b200.host` standalone.
"""
import numpy as np

from . import backend
from . import host as _host


class Compute(object):
    """the computational core: one method per reference launch site"""

    def __init__(self, ctx=None, hsfunc=None, verbose=True):
        # constructed before the parameter file is read (helios.py:40): must not depend on quant
        self._ctx = ctx
        self.hsfunc = hsfunc if hsfunc is not None else _host
        self.verbose = verbose
        self._species_cache = {}
        self._h2o_scat_lay = self._h2o_scat_int = None  # H2O Rayleigh scratch (never a resident species table)
        self._abort_sum = None
        self.fuse_passes = True
        # non-isothermal layers: form the Planck-independent part of the sweep constants once per opacity refresh
        # (helios_fband_noniso_plan_build) and run the planned sweep in the iterations in between
        self.use_flux_plan = True
        self.use_iso_plan = True  # the same for isothermal layers (helios_fband_iso_plan_build)
        # convective adjustment, convective marking + radiative-equilibrium test and the on-the-fly
        # VMR / mean-molecular-mass interpolation as device kernels (csrc/convect.cu) instead of the host functions with
        # their ~12 transfers per iteration.  Used when the host module is helios_b200.host; with the reference's
        # unchanged host_functions.py (drop-in) its own host path runs, as the drop-in contract asks.
        self.device_convection = hsfunc is None
        self.device_vmr = hsfunc is None
        self.stats = {"iterations": 0}
        backend.lib()  # fail loudly right here if the CUDA library is missing

    @property
    def ctx(self):
        if self._ctx is None:
            from . import runtime
            self._ctx = runtime.default_context()
        return self._ctx

    def _say(self, *a):
        if self.verbose:
            print(*a)

    # ------------------------------------------------------------------ set-up (C:39-82)
    def construct_planck_table(self, quant):
        self.ctx.call("plancktable", quant.dev_planckband_grid, quant.dev_opac_interwave, quant.dev_opac_deltawave,
                      quant.nbin, quant.T_star, quant.plancktable_dim, quant.plancktable_step)

    def correct_incident_energy(self, quant):
        if quant.energy_correction == 1 and quant.T_star > 10:
            import ctypes
            corr = ctypes.c_double()
            self.ctx.call("corr_inc_energy", quant.dev_planckband_grid, quant.dev_starflux, quant.dev_opac_deltawave,
                          quant.real_star, quant.nbin, quant.T_star, quant.plancktable_dim, ctypes.byref(corr))
            c = corr.value
            # the reference prints this from device code (kernels.cu:455-456)
            if c > 1:
                self._say("\nEnergy budget corrected (increased) by %.2f percent." % (100.0 * (c - 1.0)))
            if c < 1:
                self._say("\nEnergy budget corrected (decreased) by %.2f percent." % (100.0 * (1.0 - c)))

    # ------------------------------------------------------------------ per-iteration wrappers
    def interpolate_temperatures(self, quant):  # C:104-117
        self.ctx.call("temp_inter", quant.dev_T_lay, quant.dev_T_int, quant.ninterface)

    def interpolate_opacities_and_scattering_cross_sections(self, quant):  # C:119-161
        self.ctx.call("opac_interpol", quant.dev_T_lay, quant.dev_ktemp, quant.dev_p_lay, quant.dev_kpress,
                      quant.dev_opac_k, quant.dev_opac_wg_lay, quant.dev_opac_scat_cross, quant.dev_scat_cross_lay,
                      quant.npress, quant.ntemp, quant.ny, quant.nbin, quant.nlayer)
        if quant.iso == 0:
            self.ctx.call("opac_interpol", quant.dev_T_int, quant.dev_ktemp, quant.dev_p_int, quant.dev_kpress,
                          quant.dev_opac_k, quant.dev_opac_wg_int, quant.dev_opac_scat_cross, quant.dev_scat_cross_int,
                          quant.npress, quant.ntemp, quant.ny, quant.nbin, quant.ninterface)

    def interpolate_meanmolmass(self, quant):  # C:163-197
        self.ctx.call("meanmolmass_interpol", quant.dev_T_lay, quant.dev_ktemp, quant.dev_meanmolmass_lay,
                      quant.dev_opac_meanmass, quant.dev_p_lay, quant.dev_kpress, quant.npress, quant.ntemp, quant.nlayer)
        if quant.iso == 0:
            self.ctx.call("meanmolmass_interpol", quant.dev_T_int, quant.dev_ktemp, quant.dev_meanmolmass_int,
                          quant.dev_opac_meanmass, quant.dev_p_int, quant.dev_kpress, quant.npress, quant.ntemp,
                          quant.ninterface)

    @staticmethod
    def _kappa_from_file(quant):
        return isinstance(quant.input_kappa_value, str)  # C:202

    def interpolate_kappa_and_cp(self, quant):  # C:199-250
        if not self._kappa_from_file(quant):
            return
        self.ctx.call("kappa_interpol", quant.dev_T_lay, quant.dev_entr_temp, quant.dev_p_lay, quant.dev_entr_press,
                      quant.dev_kappa_lay, quant.dev_entr_kappa, quant.entr_npress, quant.entr_ntemp, quant.nlayer)
        self.ctx.call("cp_interpol", quant.dev_T_lay, quant.dev_entr_temp, quant.dev_p_lay, quant.dev_entr_press,
                      quant.dev_c_p_lay, quant.dev_entr_c_p, quant.entr_npress, quant.entr_ntemp, quant.nlayer)
        if quant.iso == 0:
            self.ctx.call("kappa_interpol", quant.dev_T_int, quant.dev_entr_temp, quant.dev_p_int,
                          quant.dev_entr_press, quant.dev_kappa_int, quant.dev_entr_kappa, quant.entr_npress,
                          quant.entr_ntemp, quant.ninterface)

    def interpolate_entropy(self, quant):  # C:252-271
        if self._kappa_from_file(quant):
            self.ctx.call("entropy_interpol", quant.dev_T_lay, quant.dev_entr_temp, quant.dev_p_lay,
                          quant.dev_entr_press, quant.dev_entropy_lay, quant.dev_entr_entropy, quant.entr_npress,
                          quant.entr_ntemp, quant.nlayer)

    def interpolate_phase_state(self, quant):  # C:273-292
        if quant.input_kappa_value == "water_atmo":
            self.ctx.call("phase_number_interpol", quant.dev_T_lay, quant.dev_entr_temp, quant.dev_p_lay,
                          quant.dev_entr_press, quant.dev_phase_number_lay, quant.dev_entr_phase_number,
                          quant.entr_npress, quant.entr_ntemp, quant.nlayer)

    def interpolate_planck(self, quant):  # C:294-329
        self.ctx.call("planck_interpol_layer", quant.dev_T_lay, quant.dev_planckband_lay, quant.dev_planckband_grid,
                      quant.dev_starflux, quant.real_star, quant.nlayer, quant.nbin, quant.plancktable_dim,
                      quant.plancktable_step)
        if quant.iso == 0:
            self.ctx.call("planck_interpol_interface", quant.dev_T_int, quant.dev_planckband_int,
                          quant.dev_planckband_grid, quant.ninterface, quant.nbin, quant.plancktable_dim,
                          quant.plancktable_step)

    def prepare_iteration(self, quant):
        """interpolate_temperatures + interpolate_planck (C:856-857) as ONE launch (helios_iteration_prepare):
        bitwise the same results; what every RT iteration does before its flux solve"""
        q = quant
        self.ctx.call("iteration_prepare", q.dev_T_lay, q.dev_T_int, q.dev_planckband_lay,
                      q.dev_planckband_int if q.iso == 0 else None, q.dev_planckband_grid, q.dev_starflux,
                      q.real_star, q.nlayer, q.nbin, q.plancktable_dim, q.plancktable_step)

    def calc_total_g_0_of_gas_and_clouds(self, quant):  # C:331-362
        self.ctx.call("calc_total_g_0_of_gas_and_clouds", quant.dev_scat_cross_lay, quant.dev_g_0_all_clouds_lay,
                      quant.dev_scat_cross_all_clouds_lay, quant.dev_g_0_tot_lay, quant.g_0, quant.nbin, quant.nlayer)
        if quant.iso == 0:
            self.ctx.call("calc_total_g_0_of_gas_and_clouds", quant.dev_scat_cross_int, quant.dev_g_0_all_clouds_int,
                          quant.dev_scat_cross_all_clouds_int, quant.dev_g_0_tot_int, quant.g_0, quant.nbin,
                          quant.ninterface)

    def calculate_transmission(self, quant):  # C:364-462
        quant._flux_plan_valid = False
        quant.dev_scat_trigger.fill_zero()  # the reference re-uploads a host zero array (C:368)
        q = quant
        tail = (q.g_0, q.epsi, q.epsi2, q.mu_star, q.w_0_limit, q.w_0_scat_limit, q.scat, q.nbin, q.ny, q.nlayer,
                q.clouds, q.scat_corr, q.debug, q.i2s_transition)
        if quant.iso == 1:
            self.ctx.call("calc_trans_iso", q.dev_trans_wg, q.dev_delta_tau_wg, q.dev_M_term, q.dev_N_term,
                          q.dev_P_term, q.dev_G_plus, q.dev_G_minus, q.dev_delta_colmass, q.dev_opac_wg_lay,
                          q.dev_meanmolmass_lay, q.dev_scat_cross_lay, q.dev_abs_cross_all_clouds_lay,
                          q.dev_scat_cross_all_clouds_lay, q.dev_delta_tau_all_clouds, q.dev_w_0, q.dev_g_0_tot_lay,
                          q.dev_scat_trigger, *tail)
        elif quant.iso == 0:
            self.ctx.call("calc_trans_noniso", q.dev_trans_wg_upper, q.dev_trans_wg_lower, q.dev_delta_tau_wg_upper,
                          q.dev_delta_tau_wg_lower, q.dev_M_upper, q.dev_M_lower, q.dev_N_upper, q.dev_N_lower,
                          q.dev_P_upper, q.dev_P_lower, q.dev_G_plus_upper, q.dev_G_plus_lower, q.dev_G_minus_upper,
                          q.dev_G_minus_lower, q.dev_delta_col_upper, q.dev_delta_col_lower, q.dev_opac_wg_lay,
                          q.dev_opac_wg_int, q.dev_meanmolmass_lay, q.dev_meanmolmass_int, q.dev_scat_cross_lay,
                          q.dev_scat_cross_int, q.dev_abs_cross_all_clouds_lay, q.dev_abs_cross_all_clouds_int,
                          q.dev_scat_cross_all_clouds_lay, q.dev_scat_cross_all_clouds_int,
                          q.dev_delta_tau_all_clouds_upper, q.dev_delta_tau_all_clouds_lower, q.dev_w_0_upper,
                          q.dev_w_0_lower, q.dev_g_0_tot_lay, q.dev_g_0_tot_int, q.dev_scat_trigger, *tail)

    def calculate_delta_z(self, quant):  # C:464-479
        self.ctx.call("calc_delta_z", quant.dev_T_lay, quant.dev_p_int, quant.dev_p_lay, quant.dev_meanmolmass_lay,
                      quant.dev_delta_z_lay, quant.g, quant.nlayer)

    def calculate_direct_beamflux(self, quant):  # C:481-526
        q = quant
        q._flux_plan_valid = False
        tail = (q.dev_z_lay, q.mu_star, q.R_planet, q.R_star, q.a, q.dir_beam, q.geom_zenith_corr, q.ninterface,
                q.nbin, q.ny)
        if quant.iso == 1:
            self.ctx.call("fdir_iso", q.dev_F_dir_wg, q.dev_planckband_lay, q.dev_delta_tau_wg, *tail)
        elif quant.iso == 0:
            self.ctx.call("fdir_noniso", q.dev_F_dir_wg, q.dev_Fc_dir_wg, q.dev_planckband_lay,
                          q.dev_delta_tau_wg_upper, q.dev_delta_tau_wg_lower, *tail)

    def build_flux_plan(self, quant):
        """B200-side addition: the sweep plan of the non-isothermal flux solve (include/helios_b200.h).  Valid until
        the next calculate_transmission / calculate_direct_beamflux."""
        q = quant
        if not self.use_flux_plan or int(q.nlayer) > (256 if q.iso == 1 else 128):
            return
        import ctypes
        if q.iso == 1:
            # isothermal layers: [a, b, Planck factor (, beam sources)] per cell
            nd = ctypes.c_size_t()
            backend._check(backend.lib().helios_fband_iso_plan_size(self.ctx.handle, int(q.ninterface), int(q.nbin),
                                                                    int(q.ny), ctypes.byref(nd)), "fband_iso_plan_size")
            n = int(nd.value)
            if n == 0 or not self.use_iso_plan:
                return
            if getattr(q, "dev_fband_plan", None) is None or q.dev_fband_plan.size != n:
                q.dev_fband_plan = self.ctx.zeros(n)
            self.ctx.call("fband_iso_plan_build", q.dev_fband_plan, q.dev_F_dir_wg, q.dev_w_0, q.dev_M_term,
                          q.dev_N_term, q.dev_P_term, q.dev_G_plus, q.dev_G_minus, q.dev_surf_albedo, q.dev_g_0_tot_lay,
                          q.g_0, q.ninterface, q.nbin, q.mu_star, q.ny, q.epsi, q.dir_beam, q.clouds, q.scat_corr,
                          q.i2s_transition)
            q._flux_plan_valid = True
            return
        nd = ctypes.c_size_t()
        backend._check(backend.lib().helios_fband_noniso_plan_size(self.ctx.handle, int(q.ninterface), int(q.nbin),
                                                                   int(q.ny), ctypes.byref(nd)), "fband_noniso_plan_size")
        n = int(nd.value)
        if n == 0:
            return
        if getattr(q, "dev_fband_plan", None) is None or q.dev_fband_plan.size != n:
            q.dev_fband_plan = self.ctx.zeros(n)
        self.ctx.call("fband_noniso_plan_build", q.dev_fband_plan, q.dev_F_dir_wg, q.dev_Fc_dir_wg, q.dev_w_0_upper,
                      q.dev_w_0_lower, q.dev_delta_tau_wg_upper, q.dev_delta_tau_wg_lower,
                      q.dev_delta_tau_all_clouds_upper, q.dev_delta_tau_all_clouds_lower, q.dev_M_upper, q.dev_M_lower,
                      q.dev_N_upper, q.dev_N_lower, q.dev_P_upper, q.dev_P_lower, q.dev_G_plus_upper,
                      q.dev_G_plus_lower, q.dev_G_minus_upper, q.dev_G_minus_lower, q.dev_surf_albedo,
                      q.dev_g_0_tot_lay, q.dev_g_0_tot_int, q.g_0, q.ninterface, q.nbin, q.mu_star, q.ny, q.epsi,
                      q.delta_tau_limit, q.clouds, q.scat_corr, q.i2s_transition)
        q._flux_plan_valid = True

    @staticmethod
    def n_scat_passes(quant):
        """number of fband launches the reference makes per call (C:531-537)"""
        nscat_step = 3 if quant.singlewalk == 0 else 1000
        return nscat_step * int(quant.scat) + 1

    def populate_spectral_flux_iteratively(self, quant):  # C:528-623
        q = quant
        npass = self.n_scat_passes(q)
        chunks = [npass] if self.fuse_passes else [1] * npass
        for n in chunks:
            if q.iso == 1 and self.use_flux_plan and self.use_iso_plan and getattr(q, "_flux_plan_valid", False):
                self.ctx.call("fband_iso_planned", q.dev_F_down_wg, q.dev_F_up_wg, q.dev_fband_plan,
                              q.dev_planckband_lay, q.dev_surf_albedo, q.R_star, q.a, q.ninterface, q.nbin, q.f_factor,
                              q.ny, q.dir_beam, n)
            elif q.iso == 1:
                self.ctx.call("fband_iso", q.dev_F_down_wg, q.dev_F_up_wg, q.dev_F_dir_wg, q.dev_planckband_lay,
                              q.dev_w_0, q.dev_M_term, q.dev_N_term, q.dev_P_term, q.dev_G_plus, q.dev_G_minus,
                              q.dev_surf_albedo, q.dev_g_0_tot_lay, q.g_0, q.singlewalk, q.R_star, q.a, q.ninterface,
                              q.nbin, q.f_factor, q.mu_star, q.ny, q.epsi, q.dir_beam, q.clouds, q.scat_corr, q.debug,
                              q.i2s_transition, n)
            elif q.iso == 0 and self.use_flux_plan and getattr(q, "_flux_plan_valid", False):
                self.ctx.call("fband_noniso_planned", q.dev_F_down_wg, q.dev_F_up_wg, q.dev_Fc_down_wg, q.dev_Fc_up_wg,
                              q.dev_fband_plan, q.dev_planckband_lay, q.dev_planckband_int, q.dev_surf_albedo,
                              q.R_star, q.a, q.ninterface, q.nbin, q.f_factor, q.ny, q.dir_beam, n)
            elif q.iso == 0:
                self.ctx.call("fband_noniso", q.dev_F_down_wg, q.dev_F_up_wg, q.dev_Fc_down_wg, q.dev_Fc_up_wg,
                              q.dev_F_dir_wg, q.dev_Fc_dir_wg, q.dev_planckband_lay, q.dev_planckband_int,
                              q.dev_w_0_upper, q.dev_w_0_lower, q.dev_delta_tau_wg_upper, q.dev_delta_tau_wg_lower,
                              q.dev_delta_tau_all_clouds_upper, q.dev_delta_tau_all_clouds_lower, q.dev_M_upper,
                              q.dev_M_lower, q.dev_N_upper, q.dev_N_lower, q.dev_P_upper, q.dev_P_lower,
                              q.dev_G_plus_upper, q.dev_G_plus_lower, q.dev_G_minus_upper, q.dev_G_minus_lower,
                              q.dev_surf_albedo, q.dev_g_0_tot_lay, q.dev_g_0_tot_int, q.g_0, q.singlewalk, q.R_star,
                              q.a, q.ninterface, q.nbin, q.f_factor, q.mu_star, q.ny, q.epsi, q.delta_tau_limit,
                              q.dir_beam, q.clouds, q.scat_corr, q.debug, q.i2s_transition, n)

    def solve_for_spectral_fluxes_via_matrix(self, quant):  # C:625-729
        q = quant
        if q.iso == 1:
            self.ctx.call("fband_matrix_iso", q.dev_F_down_wg, q.dev_F_up_wg, q.dev_F_dir_wg, q.dev_planckband_lay,
                          q.dev_w_0, q.dev_M_term, q.dev_N_term, q.dev_P_term, q.dev_G_plus, q.dev_G_minus,
                          q.dev_g_0_tot_lay, q.dev_alpha, q.dev_beta, q.dev_source_term_down, q.dev_source_term_up,
                          q.dev_c_prime, q.dev_d_prime, q.dev_scat_trigger, q.dev_trans_wg, q.dev_surf_albedo, q.g_0,
                          q.singlewalk, q.R_star, q.a, q.ninterface, q.nbin, q.f_factor, q.mu_star, q.ny, q.epsi,
                          q.dir_beam, q.clouds, q.scat_corr, q.debug, q.i2s_transition)
        if q.iso == 0:
            self.ctx.call("fband_matrix_noniso", q.dev_F_down_wg, q.dev_F_up_wg, q.dev_Fc_down_wg, q.dev_Fc_up_wg,
                          q.dev_F_dir_wg, q.dev_Fc_dir_wg, q.dev_planckband_lay, q.dev_planckband_int,
                          q.dev_w_0_upper, q.dev_w_0_lower, q.dev_delta_tau_wg_upper, q.dev_delta_tau_wg_lower,
                          q.dev_delta_tau_all_clouds_upper, q.dev_delta_tau_all_clouds_lower, q.dev_M_upper,
                          q.dev_M_lower, q.dev_N_upper, q.dev_N_lower, q.dev_P_upper, q.dev_P_lower,
                          q.dev_G_plus_upper, q.dev_G_plus_lower, q.dev_G_minus_upper, q.dev_G_minus_lower,
                          q.dev_g_0_tot_lay, q.dev_g_0_tot_int, q.dev_alpha, q.dev_beta, q.dev_source_term_down,
                          q.dev_source_term_up, q.dev_c_prime, q.dev_d_prime, q.dev_scat_trigger,
                          q.dev_trans_wg_upper, q.dev_trans_wg_lower, q.dev_surf_albedo, q.g_0, q.singlewalk,
                          q.R_star, q.a, q.ninterface, q.nbin, q.f_factor, q.mu_star, q.ny, q.epsi,
                          q.delta_tau_limit, q.dir_beam, q.clouds, q.scat_corr, q.debug, q.i2s_transition)

    def integrate_flux(self, quant):  # C:731-757
        if quant.prec == "single":
            raise ValueError("helios_b200 implements `precision = double` only")
        q = quant
        self.ctx.call("integrate_flux_double", q.dev_opac_deltawave, q.dev_F_down_tot, q.dev_F_up_tot, q.dev_F_net,
                      q.dev_F_down_wg, q.dev_F_up_wg, q.dev_F_dir_wg, q.dev_F_down_band, q.dev_F_up_band,
                      q.dev_F_dir_band, q.dev_gauss_weight, q.nbin, q.ninterface, q.ny)
        if getattr(q, "flux_allreduce", None) is not None:
            q.flux_allreduce(q)  # wavelength sharding: sum the per-interface totals over ranks

    def rad_temp_iteration(self, quant):  # C:759-797
        q = quant
        self.ctx.call("rad_temp_iter", q.dev_F_down_tot, q.dev_F_up_tot, q.dev_F_net, q.dev_F_net_diff, q.dev_T_lay,
                      q.dev_p_lay, q.dev_T_int, q.dev_p_int, q.dev_abort, q.dev_T_store, q.dev_delta_t_prefactor,
                      q.dev_F_add_heat_lay, q.dev_F_add_heat_sum, q.dev_F_smooth, q.dev_F_smooth_sum, q.dev_c_p_lay,
                      q.dev_meanmolmass_lay, q.iter_value, q.f_factor, q.foreplay, q.g, q.nlayer, q.physical_tstep,
                      q.rad_convergence_limit, q.adapt_interval, q.smooth, q.plancktable_dim, q.plancktable_step,
                      q.F_intern, q.no_atmo_mode)

    def conv_temp_iteration(self, quant):  # C:799-825
        q = quant
        self.ctx.call("conv_temp_iter", q.dev_F_down_tot, q.dev_F_up_tot, q.dev_F_net, q.dev_F_net_diff, q.dev_T_lay,
                      q.dev_p_lay, q.dev_p_int, q.dev_T_store, q.dev_delta_t_prefactor, q.dev_marked_red,
                      q.dev_F_add_heat_lay, q.dev_F_smooth, q.dev_F_smooth_sum, q.nlayer, q.iter_value,
                      q.adapt_interval, q.smooth, q.F_intern)

    # ------------------------------------------------------------------ shared pieces of the two loops
    def _refresh_atmosphere(self, quant):
        """the every-10th-iteration block (C:860-879 / C:1068-1085)"""
        hs = self.hsfunc
        if quant.opacity_mixing == "premixed":
            self.interpolate_opacities_and_scattering_cross_sections(quant)
            self.interpolate_meanmolmass(quant)
        elif quant.opacity_mixing == "on-the-fly":
            if self.device_vmr:
                self.calculate_vmr_and_meanmolmass_on_device(quant)
            else:
                hs.calculate_vmr_for_all_species(quant)
                hs.calculate_meanmolecularmass(quant)
            hs.nullify_opac_scat_arrays(quant)
            self.calculate_total_opacity_and_scat_cross_sections_from_species(quant)
        if quant.clouds == 1:
            self.calc_total_g_0_of_gas_and_clouds(quant)
        self.calculate_transmission(quant)
        self.calculate_delta_z(quant)
        quant.delta_z_lay = quant.dev_delta_z_lay.get()
        hs.calculate_height_z(quant)
        quant.dev_z_lay.set(quant.z_lay)  # in place; the reference allocates a new gpuarray (C:878)
        self.calculate_direct_beamflux(quant)
        self.build_flux_plan(quant)

    def _flux_solve(self, quant):
        if quant.flux_calc_method == "iteration":
            self.populate_spectral_flux_iteratively(quant)
        elif quant.flux_calc_method == "matrix":
            self.solve_for_spectral_fluxes_via_matrix(quant)
        else:
            print("Flux calculation method unclear. Check parameter file for typos. Aborting...")
            raise SystemExit()
        self.integrate_flux(quant)

    def _upload(self, quant, name, host):
        """`quant.dev_X = gpuarray.to_gpu(host)` without re-allocating when the size is unchanged"""
        host = np.ascontiguousarray(host)
        dev = getattr(quant, "dev_" + name, None)
        if isinstance(dev, backend.DeviceArray) and dev.size == host.size and dev.dtype == host.dtype:
            dev.set(host)
        else:
            setattr(quant, "dev_" + name, self.ctx.to_device(host))

    def _layers_converged(self, quant):
        """number of set convergence flags, reduced on the device (replaces C:927-932)"""
        if self._abort_sum is None:
            self._abort_sum = self.ctx.zeros(1, np.int32)
        self.ctx.call("abort_sum", quant.dev_abort, int(quant.nlayer) + 1, self._abort_sum)
        return int(self._abort_sum.get()[0])

    # ------------------------------------------------------------------ C:827-990
    def radiation_loop(self, quant, write, read, rt_plot):
        """loops over the relevant kernels iteratively until the equilibrium TP - profile reached"""
        hs = self.hsfunc
        keep_going = below_table_top = within_runtime = True
        quant.iter_value = np.int32(0)
        quant.p_lay = quant.dev_p_lay.get()
        quant.p_int = quant.dev_p_int.get()
        ev_total = (self.ctx.event(), self.ctx.event())
        ev_loop = (self.ctx.event(), self.ctx.event())
        time_loop = 0.0
        if quant.realtime_plot == 1:
            rt_plot.create_canvas_for_realtime_plotting()
        ev_total[0].record()
        full = int(quant.nlayer) + 1

        while keep_going and below_table_top and within_runtime:
            it = int(quant.iter_value)
            if it % 100 == 0:
                ev_loop[0].record()
            self.prepare_iteration(quant)  # interpolate_temperatures + interpolate_planck, one launch
            if it % 10 == 0:
                self._refresh_atmosphere(quant)
            self._flux_solve(quant)

            if quant.singlewalk == 1:
                break

            converged = 0
            if it < quant.foreplay:
                quant.marked_red = np.zeros(full)  # C:929: recomputed every iteration; nothing is marked before the foreplay ends
            if it % 100 == 0:
                self._say("\nWe are running \"" + str(quant.name) + "\" at iteration step nr. : " + str(it))
                if it > 99:
                    self._say("Time for the last 100 steps [s]: {:.2f}".format(time_loop * 1e-3))
            if it >= quant.foreplay:
                if quant.add_heating == 1 and it % 10 == 0:
                    hs.calc_add_heating_flux(quant)
                    self._upload(quant, "F_add_heat_lay", quant.F_add_heat_lay)
                    self._upload(quant, "F_add_heat_sum", quant.F_add_heat_sum)
                if quant.physical_tstep != 0 and it % 10 == 0:
                    self.interpolate_kappa_and_cp(quant)
                self.rad_temp_iteration(quant)
                converged = self._layers_converged(quant)
                plot_next = quant.realtime_plot == 1 and (it + 1) % quant.n_plot == 0
                if it % 100 == 0 or converged == full or plot_next:
                    # the flag array itself is only needed for reporting / plotting (realtime_plotting.py:66) / the
                    # hand-over to convection; in between the last fetched marking is kept
                    quant.abort = quant.dev_abort.get()
                    quant.marked_red = (quant.abort == 0).astype(np.float64)
                if it % 100 == 0:
                    self._say("Layers (& surface/BOA) converged: " + str(converged) + " out of " + str(full) + ".")
            keep_going = converged < full
            if quant.physical_tstep != 0:
                within_runtime = (it + 1) * quant.physical_tstep < quant.runtime_limit
            if it % 100 == 0:
                quant.T_lay = quant.dev_T_lay.get()
                below_table_top = quant.T_lay[quant.nlayer] < quant.plancktable_dim * quant.plancktable_step - 2
                if not below_table_top:
                    quant.convection = 1  # straight to the convection loop (C:949-952)

            quant.iter_value = np.int32(it + 1)
            it += 1
            if quant.realtime_plot == 1 and it % quant.n_plot == 0:
                quant.F_net = quant.dev_F_net.get()
                rt_plot.plot_tp_and_flux(quant)
            if it % 100 == 99:
                ev_loop[1].record()
                ev_loop[1].synchronize()
                time_loop = ev_loop[0].time_till(ev_loop[1])
            if quant.coupling == 1 and quant.coupl_tp_write_interval > 0:
                if it % quant.coupl_tp_write_interval == quant.coupl_tp_write_interval - 1:
                    write.write_tp_for_coupling(quant)
            if quant.crit_relaxation_numbers is not None and it in quant.crit_relaxation_numbers:
                hs.relax_radiative_convergence_criterion(quant)
            if it > quant.max_nr_iterations:
                if write is not None:
                    write.write_abort_file(quant, read)
                print("\nRun exceeds allowed maximum allowed number of iteration steps. Aborting...")
                raise SystemExit()

        ev_total[1].record()
        ev_total[1].synchronize()
        self.stats["radiation_loop_ms"] = ev_total[0].time_till(ev_total[1])
        self.stats["radiation_iterations"] = int(quant.iter_value)
        self._say("\nTime for radiative iteration [s]: {:.2f}".format(self.stats["radiation_loop_ms"] * 1e-3))
        self._say("Total number of iterative steps: " + str(quant.iter_value))

    # ------------------------------------------------------------------ C:992-1174
    def _fetch_kappa(self, quant):
        self.interpolate_kappa_and_cp(quant)
        quant.kappa_lay = quant.dev_kappa_lay.get()
        if quant.iso == 0:
            quant.kappa_int = quant.dev_kappa_int.get()

    def _pull_convection_state(self, quant):
        """host copies of what the output writers and the feedback prints read after / during the device-side loop"""
        q = quant
        for name in ("T_lay", "F_net", "F_down_tot", "F_up_tot", "F_net_diff", "F_smooth_sum", "conv_layer", "marked_red"):
            setattr(q, name, getattr(q, "dev_" + name).get())

    def convection_loop(self, quant, write, read, rt_plot):
        """loops interchangeably through the radiative and convection schemes"""
        if not (quant.singlewalk == 0 and quant.convection == 1):
            return
        hs = self.hsfunc
        self._fetch_kappa(quant)
        quant.T_lay = quant.dev_T_lay.get()
        quant.p_lay = quant.dev_p_lay.get()
        quant.p_int = quant.dev_p_int.get()
        if quant.iso == 0:
            hs.conv_check(quant)
            hs.mark_convective_layers(quant, stitching=0)
        if quant.conv_unstable is None:
            # isothermal layers: the reference never runs its stability check here and then fails on
            # sum(None) (C:1004-1009); there is no interface kappa to check against
            raise ValueError("convective adjustment needs non-isothermal layers (iso = 0), as in the reference")
        unstable = sum(quant.conv_unstable) > 0

        ev_total = (self.ctx.event(), self.ctx.event())
        ev_loop = (self.ctx.event(), self.ctx.event())
        time_loop = 0.0
        ev_total[0].record()
        quant.iter_value = np.int32(0)
        if unstable:
            self._say("\nConvectively unstable layers found. Starting convective adjustment")
        else:
            self._say("\nAll layers convectively stable. No convective adjustment necessary.\n")
        quant.F_net = quant.dev_F_net.get()
        quant.F_up_tot = quant.dev_F_up_tot.get()
        quant.F_down_tot = quant.dev_F_down_tot.get()

        if self.device_convection and quant.iso == 0:
            self._convection_buffers(quant)
            quant.dev_conv_layer.set(np.asarray(quant.conv_layer, np.int32))
        running = unstable
        try:
            self._convection_iterations(quant, write, read, rt_plot, running, ev_loop)
        finally:
            if self.device_convection and quant.iso == 0 and unstable:
                self._pull_convection_state(quant)  # also when the iteration guard ends the run (C:1157-1164)
        ev_total[1].record()
        ev_total[1].synchronize()
        self.stats["convection_loop_ms"] = ev_total[0].time_till(ev_total[1])
        self.stats["convection_iterations"] = int(quant.iter_value)
        self._say("\nTime for rad.-conv. iteration [s]: {:.2f}".format(self.stats["convection_loop_ms"] * 1e-3))
        self._say("Total number of iterative steps: " + str(quant.iter_value))

    def _convection_iterations(self, quant, write, read, rt_plot, running, ev_loop):
        """the body of the radiative-convective loop (C:1030-1164)"""
        hs = self.hsfunc
        time_loop = 0.0
        while running:
            it = int(quant.iter_value)
            if it % 100 == 0:
                ev_loop[0].record()
                self._say("\nWe are running \"" + str(quant.name) + "\" at iteration step nr. : " + str(it))
                if it > 99:
                    self._say("Time for the last 100 steps [s]: {:.2f}".format(time_loop * 1e-3))

            on_device = self.device_convection and quant.iso == 0
            # 1. convective adjustment
            self.interpolate_temperatures(quant)
            if it % 10 == 0:
                if quant.opacity_mixing == "premixed":
                    self.interpolate_meanmolmass(quant)
                elif quant.opacity_mixing == "on-the-fly":
                    if self.device_vmr:
                        self.calculate_vmr_and_meanmolmass_on_device(quant)
                    else:
                        hs.calculate_vmr_for_all_species(quant)
                        hs.calculate_meanmolecularmass(quant)
            if on_device:
                # one launch on resident arrays instead of six transfers and the host loops of H:337-538
                self.interpolate_kappa_and_cp(quant)
                self.convective_adjustment_on_device(quant)
            else:
                self._fetch_kappa(quant)
                quant.c_p_lay = quant.dev_c_p_lay.get()
                quant.meanmolmass_lay = quant.dev_meanmolmass_lay.get()
                quant.T_lay = quant.dev_T_lay.get()
                quant.F_smooth_sum = quant.dev_F_smooth_sum.get()
                hs.convective_adjustment(quant)
                quant.dev_T_lay.set(quant.T_lay)

            # 2. radiative fluxes for the adjusted profile
            self.interpolate_temperatures(quant)
            self.interpolate_planck(quant)
            if it % 10 == 0:
                self._refresh_atmosphere(quant)
            self._flux_solve(quant)

            # 3. mark the convective zones for the convergence test
            if on_device:
                self.interpolate_kappa_and_cp(quant)
                n_conv_ok, n_rad, n_conv, zero_T = self.convection_marks_on_device(quant)
                if zero_T:
                    print("WARNING WARNING WARNING: Found zero temperature in", zero_T, "layer(s)")
                if quant.physical_tstep != 0:
                    break
                if it % 100 == 1:
                    self._say("Number of radiative layers converged: {:d} out of {:d}.".format(n_conv_ok, n_rad))
                running = (n_conv_ok != n_rad) or (it < 400) or (n_conv == 0)
                if it % 100 == 1 and self.verbose:
                    self._pull_convection_state(quant)
                    hs.give_feedback_on_convergence(quant)
            else:
                quant.F_net = quant.dev_F_net.get()
                quant.F_down_tot = quant.dev_F_down_tot.get()
                quant.F_up_tot = quant.dev_F_up_tot.get()
                quant.F_net_diff = quant.dev_F_net_diff.get()
                self._fetch_kappa(quant)
                quant.T_lay = quant.dev_T_lay.get()
                hs.mark_convective_layers(quant, stitching=1)
                if quant.physical_tstep != 0:
                    break
                quant.F_smooth_sum = quant.dev_F_smooth_sum.get()
                running = (not hs.check_for_radiative_eq(quant)) or (it < 400) or (sum(quant.conv_layer) == 0)
                if it % 100 == 1 and self.verbose:
                    hs.give_feedback_on_convergence(quant)

            # 4. radiative forward step where the local criterion is not met
            if running:
                if quant.realtime_plot == 1 and it % quant.n_plot == 0:
                    rt_plot.plot_tp_and_flux(quant)
                if quant.add_heating == 1 and it % 10 == 0:
                    hs.calc_add_heating_flux(quant)
                    self._upload(quant, "F_add_heat_lay", quant.F_add_heat_lay)
                    self._upload(quant, "F_add_heat_sum", quant.F_add_heat_sum)
                if not on_device:
                    self._upload(quant, "conv_layer", np.asarray(quant.conv_layer, np.int32))
                    self._upload(quant, "marked_red", np.asarray(quant.marked_red, np.int32))
                self.conv_temp_iteration(quant)
                if not on_device:
                    quant.T_lay = quant.dev_T_lay.get()
                if it % 100 == 99:
                    ev_loop[1].record()
                    ev_loop[1].synchronize()
                    time_loop = ev_loop[0].time_till(ev_loop[1])
                quant.iter_value = np.int32(it + 1)
                it += 1

            if quant.coupling == 1 and quant.coupl_tp_write_interval > 0:
                if it % quant.coupl_tp_write_interval == quant.coupl_tp_write_interval - 1:
                    write.write_tp_for_coupling(quant)
            if quant.crit_relaxation_numbers is not None and it in quant.crit_relaxation_numbers:
                hs.relax_radiative_convergence_criterion(quant)
            if it > quant.max_nr_iterations:
                if write is not None:
                    write.write_abort_file(quant, read)
                print("\nRun exceeds allowed maximum allowed number of iteration steps. Aborting...")
                raise SystemExit()


    # ------------------------------------------------------------------ post-processing (C:1176-1296)
    def integrate_optdepth_transmission(self, quant):
        q = quant
        if q.iso == 1:
            self.ctx.call("integrate_optdepth_transmission_iso", q.dev_trans_wg, q.dev_trans_band, q.dev_delta_tau_wg,
                          q.dev_delta_tau_band, q.dev_gauss_weight, q.nbin, q.nlayer, q.ny)
        elif q.iso == 0:
            self.ctx.call("integrate_optdepth_transmission_noniso", q.dev_trans_wg_upper, q.dev_trans_wg_lower,
                          q.dev_trans_band, q.dev_delta_tau_wg_upper, q.dev_delta_tau_wg_lower, q.dev_delta_tau_band,
                          q.dev_gauss_weight, q.dev_delta_tau_all_clouds, q.dev_delta_tau_all_clouds_upper,
                          q.dev_delta_tau_all_clouds_lower, q.nbin, q.nlayer, q.ny)

    def calculate_contribution_function(self, quant):
        q = quant
        if q.iso == 1:
            self.ctx.call("calc_contr_func_iso", q.dev_trans_wg, q.dev_trans_weight_band, q.dev_contr_func_band,
                          q.dev_gauss_weight, q.dev_planckband_lay, q.epsi, q.nbin, q.nlayer, q.ny)
        elif q.iso == 0:
            self.ctx.call("calc_contr_func_noniso", q.dev_trans_wg_upper, q.dev_trans_wg_lower,
                          q.dev_trans_weight_band, q.dev_contr_func_band, q.dev_gauss_weight, q.dev_planckband_lay,
                          q.epsi, q.nbin, q.nlayer, q.ny)

    def calculate_mean_opacities(self, quant):
        q = quant
        self.ctx.call("calc_mean_opacities", q.dev_planck_opac_T_pl, q.dev_ross_opac_T_pl, q.dev_planck_opac_T_star,
                      q.dev_ross_opac_T_star, q.dev_opac_wg_lay, q.dev_abs_cross_all_clouds_lay, q.dev_meanmolmass_lay,
                      q.dev_planckband_lay, q.dev_opac_interwave, q.dev_opac_deltawave, q.dev_T_lay,
                      q.dev_gauss_weight, q.dev_gauss_y, q.dev_opac_band_lay, q.nlayer, q.nbin, q.ny, q.T_star)

    def integrate_beamflux(self, quant):
        q = quant
        self.ctx.call("integrate_beamflux", q.dev_F_dir_tot, q.dev_F_dir_band, q.dev_opac_deltawave,
                      q.dev_gauss_weight, q.nbin, q.ninterface)

    # ------------------------------------------------------------------ on-the-fly mixing (C:1298-1501)
    def interpolate_species_opac(self, quant):
        q = quant
        self.ctx.call("opac_species_interpol", q.dev_T_lay, q.dev_ktemp, q.dev_p_lay, q.dev_kpress,
                      q.dev_opacity_spec_pretab, q.dev_opac_spec_wg_lay, q.npress, q.ntemp, q.ny, q.nbin, q.nlayer)
        if q.iso == 0:
            self.ctx.call("opac_species_interpol", q.dev_T_int, q.dev_ktemp, q.dev_p_int, q.dev_kpress,
                          q.dev_opacity_spec_pretab, q.dev_opac_spec_wg_int, q.npress, q.ntemp, q.ny, q.nbin,
                          q.ninterface)

    def add_to_mixed_opacity(self, quant, mass_spec, s):
        q = quant
        mass_spec = np.float64(mass_spec * _amu(self.hsfunc))
        ro_method = 0 if (q.kcoeff_mixing == "correlated-k" or "CIA" in q.species_list[s].name) else 1
        self.ctx.call("add_to_mixed_opac", q.dev_vmr_spec_lay, q.dev_opac_spec_wg_lay, q.dev_opac_wg_lay,
                      q.dev_meanmolmass_lay, q.dev_gauss_weight, q.dev_gauss_y, mass_spec, s, ro_method, q.ny, q.nbin,
                      q.nlayer)
        if q.iso == 0:
            self.ctx.call("add_to_mixed_opac", q.dev_vmr_spec_int, q.dev_opac_spec_wg_int, q.dev_opac_wg_int,
                          q.dev_meanmolmass_int, q.dev_gauss_weight, q.dev_gauss_y, mass_spec, s, ro_method, q.ny,
                          q.nbin, q.ninterface)

    def calculate_H2O_Rayleigh_scattering(self, quant, s):
        q = quant
        mass_h2o = np.float64(q.species_list[s].weight * _amu(self.hsfunc))
        self.ctx.call("calc_h2o_scat", q.dev_T_lay, q.dev_p_lay, q.dev_opac_wave, q.dev_scat_cross_spec_lay,
                      q.dev_vmr_spec_lay, mass_h2o, q.nbin, q.nlayer)
        if q.iso == 0:
            self.ctx.call("calc_h2o_scat", q.dev_T_int, q.dev_p_int, q.dev_opac_wave, q.dev_scat_cross_spec_int,
                          q.dev_vmr_spec_int, mass_h2o, q.nbin, q.ninterface)

    def add_to_mixed_scat_cross_sect(self, quant):
        q = quant
        self.ctx.call("add_to_mixed_scat", q.dev_vmr_spec_lay, q.dev_scat_cross_spec_lay, q.dev_scat_cross_lay, q.nbin,
                      q.nlayer)
        if q.iso == 0:
            self.ctx.call("add_to_mixed_scat", q.dev_vmr_spec_int, q.dev_scat_cross_spec_int, q.dev_scat_cross_int,
                          q.nbin, q.ninterface)

    def _resident(self, key, host_obj):
        """species tables stay in HBM: the reference re-uploads each ~200 MB table on every call (C:1467).
        Keyed by the identity of the host object; replace the object (not its contents) to refresh."""
        hit = self._species_cache.get(key)
        if hit is not None and hit[0] is host_obj:
            return hit[1]
        dev = self.ctx.to_device(np.ascontiguousarray(host_obj, dtype=np.float64))
        self._species_cache[key] = (host_obj, dev)
        return dev

    def calculate_total_opacity_and_scat_cross_sections_from_species(self, quant):
        q = quant
        on_dev = getattr(q, "_vmr_on_device", None)
        for s, sp in enumerate(q.species_list):
            if on_dev is not None:
                q.dev_vmr_spec_lay = on_dev[s][0]  # interpolated on the device: nothing to upload (C:1461-1463)
                if q.iso == 0:
                    q.dev_vmr_spec_int = on_dev[s][1]
            else:
                self._upload(q, "vmr_spec_lay", np.asarray(sp.vmr_layer, np.float64))
                if q.iso == 0:
                    self._upload(q, "vmr_spec_int", np.asarray(sp.vmr_interface, np.float64))
            if sp.absorbing == "yes":
                q.dev_opacity_spec_pretab = self._resident(("k", s), sp.opacity_pretab)
                self.interpolate_species_opac(q)
                self.add_to_mixed_opacity(q, sp.weight, s)
            if sp.scattering == "yes":
                if sp.name == "H2O":
                    # private scratch: calc_h2o_scat writes every element, so it must never land in another
                    # species' resident cross-section table (the reference allocates fresh zeros, C:1483-1486)
                    n_lay, n_int = int(q.nbin) * int(q.nlayer), int(q.nbin) * int(q.ninterface)
                    if self._h2o_scat_lay is None or self._h2o_scat_lay.size != n_lay:
                        self._h2o_scat_lay = self.ctx.zeros(n_lay)
                    q.dev_scat_cross_spec_lay = self._h2o_scat_lay
                    if q.iso == 0:
                        if self._h2o_scat_int is None or self._h2o_scat_int.size != n_int:
                            self._h2o_scat_int = self.ctx.zeros(n_int)
                        q.dev_scat_cross_spec_int = self._h2o_scat_int
                    self.calculate_H2O_Rayleigh_scattering(q, s)
                else:
                    q.dev_scat_cross_spec_lay = self._resident(("sl", s), sp.scat_cross_sect_layer)
                    if q.iso == 0:
                        q.dev_scat_cross_spec_int = self._resident(("si", s), sp.scat_cross_sect_interface)
                self.add_to_mixed_scat_cross_sect(q)


    # ------------------------------------------------------------------ device-side host functions (csrc/convect.cu)
    def calculate_vmr_and_meanmolmass_on_device(self, quant):
        """H:874-924 without leaving the device: per species the VMR profile (bilinear in (T, log10 P) for FastChem-tabulated
        species, the given constant profile otherwise) and the VMR-weighted mean molecular mass.  The per-species
        profiles stay resident; the species loop (C:1454-1501) then takes them from there."""
        q = quant
        nl, ni = int(q.nlayer), int(q.ninterface)
        levels = [("lay", q.dev_T_lay, q.dev_p_lay, nl)] + ([("int", q.dev_T_int, q.dev_p_int, ni)] if q.iso == 0 else [])
        cache = getattr(q, "_vmr_on_device", None)
        if cache is None or len(cache) != len(q.species_list):
            cache = [[None, None] for _ in q.species_list]
        scratch = getattr(q, "_mmm_scratch", None)
        if scratch is None:
            scratch = q._mmm_scratch = [self.ctx.zeros(ni), self.ctx.zeros(ni)]
        for k, (tag, dT, dP, n) in enumerate(levels):
            scratch[0].fill_zero()
            scratch[1].fill_zero()
            for s, sp in enumerate(q.species_list):
                if getattr(sp, "source_for_vmr", None) == "FastChem":
                    if cache[s][k] is None or cache[s][k].size != n:
                        cache[s][k] = self.ctx.zeros(n)
                    table = self._resident(("vmr", s), sp.vmr_pretab)
                    self.ctx.call("vmr_interpol", dT, dP, q.dev_ktemp, q.dev_kpress, table, cache[s][k], q.npress, q.ntemp, n)
                else:
                    host = sp.vmr_layer if tag == "lay" else sp.vmr_interface
                    cache[s][k] = self._resident(("vmr_" + tag, s), host)
                if "CIA" in sp.name or sp.name in ("H-_ff", "He-"):
                    continue  # not part of the mean molecular mass (H:941-943)
                self.ctx.call("meanmolmass_accumulate", cache[s][k], float(sp.weight), scratch[0], scratch[1], n)
            out = q.dev_meanmolmass_lay if tag == "lay" else q.dev_meanmolmass_int
            self.ctx.call("meanmolmass_finish", scratch[0], scratch[1], out, n)
        q._vmr_on_device = cache

    def _convection_buffers(self, quant):
        q = quant
        n1 = int(q.nlayer) + 1
        for name in ("conv_layer", "conv_unstable", "marked_red"):
            dev = getattr(q, "dev_" + name, None)
            if not isinstance(dev, backend.DeviceArray) or dev.size != n1 or dev.dtype != np.int32:
                setattr(q, "dev_" + name, self.ctx.zeros(n1, np.int32))
        if getattr(q, "_conv_status", None) is None:
            q._conv_status = self.ctx.zeros(8, np.int32)

    def convective_adjustment_on_device(self, quant):
        """H:509-538 in one launch; returns nothing (the status words stay on the device until asked for)"""
        q = quant
        dampara = -1.0 if q.input_dampara == "automatic" else float(q.input_dampara)
        self.ctx.call("convective_adjustment", q.dev_T_lay, q.dev_p_lay, q.dev_p_int, q.dev_kappa_lay, q.dev_kappa_int,
                      q.dev_c_p_lay, q.dev_meanmolmass_lay, q.dev_F_add_heat_sum, q.dev_F_smooth_sum, q.dev_F_down_tot,
                      q.dev_F_up_tot, q.dev_conv_layer, q.dev_conv_unstable, q._conv_status, q.F_intern, q.T_star, dampara,
                      q.iter_value, q.nlayer)

    def convection_marks_on_device(self, quant):
        """H:545-582 (stitching = 1) + H:251-286; returns (converged radiative layers, radiative layers, convective layers)"""
        q = quant
        self.ctx.call("convection_marks", q.dev_T_lay, q.dev_p_lay, q.dev_p_int, q.dev_kappa_lay, q.dev_kappa_int,
                      q.dev_F_net, q.dev_F_down_tot, q.dev_F_add_heat_sum, q.dev_F_smooth_sum, q.dev_conv_layer,
                      q.dev_marked_red, q._conv_status.view(4, 4), q.F_intern, q.rad_convergence_limit, q.iter_value,
                      q.nlayer)
        st = q._conv_status.get()
        return int(st[4]), int(st[5]), int(st[6]), int(st[7])


def _amu(hsfunc):
    pc = getattr(hsfunc, "pc", None)
    return pc.AMU if pc is not None else _host.AMU
