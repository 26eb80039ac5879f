"""CPU oracle driver: the launch sites of computation.py executed with helios_oracle.py on NumPy arrays.
TEST INFRASTRUCTURE, NOT PRODUCT (see the header of helios_oracle.py).

`HostMirror(q)` snapshots a Store: every `dev_*` DeviceArray becomes a NumPy array of the same name and
every scalar is copied.  `OracleCompute` then offers the method names of `Compute`, reading and writing
those arrays, so a parity test is: run the method on the GPU, run it on a mirror taken just before,
compare the outputs.
"""
import copy

import numpy as np

from . import helios_oracle as O


class HostMirror(object):
    def __init__(self, q):
        for k, v in vars(q).items():
            if k.startswith("_"):
                continue
            if hasattr(v, "get") and hasattr(v, "ptr"):
                setattr(self, k, v.get())
            elif isinstance(v, np.ndarray):
                setattr(self, k, v.copy())
            elif k == "species_list":
                setattr(self, k, v)
            else:
                try:
                    setattr(self, k, copy.copy(v))
                except Exception:
                    setattr(self, k, v)


def mirror_from_host(q):
    """A mirror built from a Store that has NOT been uploaded: what copy_host_to_device() and
    allocate_on_device() would have put on the GPU, as NumPy arrays.  Lets the whole oracle pipeline
    run without a GPU."""
    from helios_b200 import quantities as Q
    q.create_zero_arrays()
    q.convert_input_list_to_array()
    m = HostMirror(q)
    names = list(Q._INPUTS) + list(Q._INPUTS_NONISO)
    for name in names:
        v = getattr(q, name)
        setattr(m, "dev_" + name, None if v is None else np.array(v, np.float64))
    for name, _, kind, _ in Q._ZEROS:
        setattr(m, "dev_" + name, getattr(q, name).copy())
    otf = q.opacity_mixing == "on-the-fly"
    for name, size_key, cond in Q._DEVICE_ONLY:
        need = (cond == "always" or (cond == "noniso" and q.iso == 0) or (cond == "otf" and otf)
                or (cond == "noniso+otf" and q.iso == 0 and otf) or (cond == "matrix" and q.flux_calc_method == "matrix"))
        if need:
            setattr(m, "dev_" + name, np.zeros(q._size(size_key)))
    m.dev_marked_red = np.zeros(int(q.nlayer) + 1, np.int32)
    return m


class OracleCompute(object):
    def construct_planck_table(self, q):
        q.dev_planckband_grid = O.plancktable(q.dev_opac_interwave, q.dev_opac_deltawave, q.T_star,
                                              int(q.plancktable_dim), int(q.plancktable_step))

    def correct_incident_energy(self, q):
        if q.energy_correction == 1 and q.T_star > 10:
            g, s, c = O.corr_inc_energy(q.dev_planckband_grid, q.dev_starflux, q.dev_opac_deltawave, int(q.real_star),
                                        int(q.nbin), q.T_star, int(q.plancktable_dim))
            q.dev_planckband_grid = g
            if s is not None:
                q.dev_starflux = s
            return c

    def interpolate_temperatures(self, q):
        q.dev_T_int = O.temp_inter(q.dev_T_lay, int(q.ninterface))

    def interpolate_planck(self, q):
        q.dev_planckband_lay = O.planck_interpol_layer(q.dev_T_lay, q.dev_planckband_grid, q.dev_starflux,
                                                       int(q.real_star), int(q.nlayer), int(q.nbin),
                                                       int(q.plancktable_dim), int(q.plancktable_step))
        if q.iso == 0:
            q.dev_planckband_int = O.planck_interpol_interface(q.dev_T_int, q.dev_planckband_grid, int(q.ninterface),
                                                               int(q.nbin), int(q.plancktable_dim),
                                                               int(q.plancktable_step))

    @staticmethod
    def _pad(arr, n):
        out = np.zeros(n, arr.dtype)
        out[:arr.size] = arr
        return out

    def interpolate_opacities_and_scattering_cross_sections(self, q):
        k, s = O.opac_interpol(q.dev_T_lay, q.dev_ktemp, q.dev_p_lay, q.dev_kpress, q.dev_opac_k, q.dev_opac_scat_cross,
                               int(q.npress), int(q.ntemp), int(q.ny), int(q.nbin), int(q.nlayer))
        q.dev_opac_wg_lay = self._pad(k, q.dev_opac_wg_lay.size)
        q.dev_scat_cross_lay = s
        if q.iso == 0:
            k, s = O.opac_interpol(q.dev_T_int, q.dev_ktemp, q.dev_p_int, q.dev_kpress, q.dev_opac_k,
                                   q.dev_opac_scat_cross, int(q.npress), int(q.ntemp), int(q.ny), int(q.nbin),
                                   int(q.ninterface))
            q.dev_opac_wg_int = k
            q.dev_scat_cross_int = s

    def interpolate_meanmolmass(self, q):
        q.dev_meanmolmass_lay = O.scalar_interpol(q.dev_T_lay, q.dev_ktemp, q.dev_p_lay, q.dev_kpress,
                                                  q.dev_opac_meanmass, int(q.npress), int(q.ntemp), int(q.nlayer))
        if q.iso == 0:
            q.dev_meanmolmass_int = O.scalar_interpol(q.dev_T_int, q.dev_ktemp, q.dev_p_int, q.dev_kpress,
                                                      q.dev_opac_meanmass, int(q.npress), int(q.ntemp),
                                                      int(q.ninterface))

    def calc_total_g_0_of_gas_and_clouds(self, q):
        q.dev_g_0_tot_lay = O.calc_total_g_0(q.dev_scat_cross_lay, q.dev_g_0_all_clouds_lay,
                                             q.dev_scat_cross_all_clouds_lay, q.g_0)
        if q.iso == 0:
            q.dev_g_0_tot_int = O.calc_total_g_0(q.dev_scat_cross_int, q.dev_g_0_all_clouds_int,
                                                 q.dev_scat_cross_all_clouds_int, q.g_0)

    _ISO_MAP = dict(trans="trans_wg", dtau="delta_tau_wg", M="M_term", N="N_term", P="P_term", Gp="G_plus",
                    Gm="G_minus", w0="w_0")

    _NONISO_MAP = dict(trans="trans_wg", dtau="delta_tau_wg", M="M", N="N", P="P", Gp="G_plus", Gm="G_minus", w0="w_0")

    def calculate_transmission(self, q):
        nl, nb, ny = int(q.nlayer), int(q.nbin), int(q.ny)
        if q.iso == 1:
            r = O.calc_trans_iso(q.dev_delta_colmass, q.dev_opac_wg_lay, q.dev_meanmolmass_lay, q.dev_scat_cross_lay,
                                 q.dev_abs_cross_all_clouds_lay, q.dev_scat_cross_all_clouds_lay, q.dev_g_0_tot_lay,
                                 q.g_0, q.epsi, q.epsi2, q.mu_star, q.w_0_limit, q.w_0_scat_limit, int(q.scat), nb, ny,
                                 nl, int(q.clouds), int(q.scat_corr), q.i2s_transition)
            for k, name in self._ISO_MAP.items():
                setattr(q, "dev_" + name, self._pad(r[k], getattr(q, "dev_" + name).size))
            q.dev_delta_tau_all_clouds = r["dtau_clouds"]
            q.dev_scat_trigger = r["scat_trigger"]
        else:
            up, low, trig = O.calc_trans_noniso(
                q.dev_delta_col_upper, q.dev_delta_col_lower, q.dev_opac_wg_lay, q.dev_opac_wg_int,
                q.dev_meanmolmass_lay, q.dev_meanmolmass_int, q.dev_scat_cross_lay, q.dev_scat_cross_int,
                q.dev_abs_cross_all_clouds_lay, q.dev_abs_cross_all_clouds_int, q.dev_scat_cross_all_clouds_lay,
                q.dev_scat_cross_all_clouds_int, q.dev_g_0_tot_lay, q.dev_g_0_tot_int, q.g_0, q.epsi, q.epsi2,
                q.mu_star, q.w_0_limit, q.w_0_scat_limit, int(q.scat), nb, ny, nl, int(q.clouds), int(q.scat_corr),
                q.i2s_transition)
            for half, sfx in ((up, "_upper"), (low, "_lower")):
                for k, base in self._NONISO_MAP.items():
                    full = "dev_" + base + sfx
                    setattr(q, full, self._pad(half[k], getattr(q, full).size))
                setattr(q, "dev_delta_tau_all_clouds" + sfx, half["dtau_clouds"])
            q.dev_scat_trigger = trig

    def calculate_delta_z(self, q):
        q.dev_delta_z_lay = O.calc_delta_z(q.dev_T_lay, q.dev_p_int, q.dev_meanmolmass_lay, q.g, int(q.nlayer))

    def calculate_direct_beamflux(self, q):
        args = (q.dev_z_lay, q.mu_star, q.R_planet, q.R_star, q.a, int(q.dir_beam), int(q.geom_zenith_corr),
                int(q.ninterface), int(q.nbin), int(q.ny))
        if q.iso == 1:
            q.dev_F_dir_wg, _ = O.fdir(q.dev_planckband_lay, q.dev_delta_tau_wg, None, *args)
        else:
            F, Fc = O.fdir(q.dev_planckband_lay, q.dev_delta_tau_wg_upper, q.dev_delta_tau_wg_lower, *args)
            Fc_old = q.dev_Fc_dir_wg.reshape(int(q.ninterface), -1)
            Fc = Fc.reshape(int(q.ninterface), -1)
            Fc[-1] = Fc_old[-1]  # never written by the kernel
            q.dev_F_dir_wg, q.dev_Fc_dir_wg = F, Fc.reshape(-1)

    @staticmethod
    def _halves(q):
        def h(sfx):
            return dict(w0=getattr(q, "dev_w_0" + sfx), dtau=getattr(q, "dev_delta_tau_wg" + sfx),
                        dtau_clouds=getattr(q, "dev_delta_tau_all_clouds" + sfx), M=getattr(q, "dev_M" + sfx),
                        N=getattr(q, "dev_N" + sfx), P=getattr(q, "dev_P" + sfx), Gp=getattr(q, "dev_G_plus" + sfx),
                        Gm=getattr(q, "dev_G_minus" + sfx), trans=getattr(q, "dev_trans_wg" + sfx))
        return h("_upper"), h("_lower")

    def populate_spectral_flux_iteratively(self, q, npass=None):
        if npass is None:
            npass = (3 if q.singlewalk == 0 else 1000) * int(q.scat) + 1
        if q.iso == 1:
            q.dev_F_down_wg, q.dev_F_up_wg = O.fband_iso(
                q.dev_F_down_wg, q.dev_F_up_wg, q.dev_F_dir_wg, q.dev_planckband_lay, q.dev_w_0, q.dev_M_term,
                q.dev_N_term, q.dev_P_term, q.dev_G_plus, q.dev_G_minus, q.dev_surf_albedo, q.dev_g_0_tot_lay, q.g_0,
                q.R_star, q.a, int(q.ninterface), int(q.nbin), q.f_factor, q.mu_star, int(q.ny), q.epsi,
                int(q.dir_beam), int(q.clouds), int(q.scat_corr), q.i2s_transition, npass)
        else:
            up, low = self._halves(q)
            (q.dev_F_down_wg, q.dev_F_up_wg, q.dev_Fc_down_wg, q.dev_Fc_up_wg) = O.fband_noniso(
                q.dev_F_down_wg, q.dev_F_up_wg, q.dev_Fc_down_wg, q.dev_Fc_up_wg, q.dev_F_dir_wg, q.dev_Fc_dir_wg,
                q.dev_planckband_lay, q.dev_planckband_int, up, low, q.dev_surf_albedo, q.dev_g_0_tot_lay,
                q.dev_g_0_tot_int, q.g_0, q.R_star, q.a, int(q.ninterface), int(q.nbin), q.f_factor, q.mu_star,
                int(q.ny), q.epsi, q.delta_tau_limit, int(q.dir_beam), int(q.clouds), int(q.scat_corr),
                q.i2s_transition, npass)

    def solve_for_spectral_fluxes_via_matrix(self, q):
        if q.iso == 1:
            q.dev_F_down_wg, q.dev_F_up_wg = O.fband_matrix_iso(
                q.dev_F_down_wg, q.dev_F_up_wg, q.dev_F_dir_wg, q.dev_planckband_lay, q.dev_w_0, q.dev_M_term,
                q.dev_N_term, q.dev_P_term, q.dev_G_plus, q.dev_G_minus, q.dev_g_0_tot_lay, q.dev_scat_trigger,
                q.dev_trans_wg, q.dev_surf_albedo, q.g_0, q.R_star, q.a, int(q.ninterface), int(q.nbin), q.f_factor,
                q.mu_star, int(q.ny), q.epsi, int(q.dir_beam), int(q.clouds), int(q.scat_corr), q.i2s_transition)
        else:
            up, low = self._halves(q)
            (q.dev_F_down_wg, q.dev_F_up_wg, q.dev_Fc_down_wg, q.dev_Fc_up_wg) = O.fband_matrix_noniso(
                q.dev_F_down_wg, q.dev_F_up_wg, q.dev_Fc_down_wg, q.dev_Fc_up_wg, q.dev_F_dir_wg, q.dev_Fc_dir_wg,
                q.dev_planckband_lay, q.dev_planckband_int, up, low, q.dev_g_0_tot_lay, q.dev_g_0_tot_int,
                q.dev_scat_trigger, q.dev_surf_albedo, q.g_0, q.R_star, q.a, int(q.ninterface), int(q.nbin),
                q.f_factor, q.mu_star, int(q.ny), q.epsi, q.delta_tau_limit, int(q.dir_beam), int(q.clouds),
                int(q.scat_corr), q.i2s_transition)

    def integrate_flux(self, q):
        r = O.integrate_flux(q.dev_opac_deltawave, q.dev_F_down_wg, q.dev_F_up_wg, q.dev_F_dir_wg, q.dev_gauss_weight,
                             int(q.nbin), int(q.ninterface), int(q.ny))
        for k, v in r.items():
            setattr(q, "dev_" + k, v)

    def rad_temp_iteration(self, q):
        r = O.rad_temp_iter(q.dev_F_down_tot, q.dev_F_net, q.dev_T_lay, q.dev_p_lay, q.dev_p_int, q.dev_T_store,
                            q.dev_delta_t_prefactor, q.dev_F_add_heat_lay, q.dev_F_add_heat_sum, q.dev_F_smooth,
                            q.dev_F_smooth_sum, q.dev_c_p_lay, q.dev_meanmolmass_lay, int(q.iter_value),
                            int(q.foreplay), q.g, int(q.nlayer), q.physical_tstep, q.rad_convergence_limit,
                            int(q.adapt_interval), int(q.smooth), int(q.plancktable_dim), int(q.plancktable_step),
                            q.F_intern, int(q.no_atmo_mode))
        q.dev_T_lay, q.dev_abort, q.dev_T_store = r["tlay"], r["abrt"], r["T_store"]
        q.dev_delta_t_prefactor, q.dev_F_net_diff = r["prefactor"], r["F_net_diff"]
        q.dev_F_smooth, q.dev_F_smooth_sum = r["F_smooth"], r["F_smooth_sum"]

    def conv_temp_iteration(self, q):
        r = O.conv_temp_iter(q.dev_F_net, q.dev_T_lay, q.dev_p_lay, q.dev_p_int, q.dev_T_store,
                             q.dev_delta_t_prefactor, q.dev_marked_red, q.dev_F_add_heat_lay, q.dev_F_smooth,
                             q.dev_F_smooth_sum, int(q.nlayer), int(q.iter_value), int(q.adapt_interval),
                             int(q.smooth), q.F_intern)
        q.dev_T_lay, q.dev_T_store = r["tlay"], r["T_store"]
        q.dev_delta_t_prefactor, q.dev_F_net_diff = r["prefactor"], r["F_net_diff"]
        q.dev_F_smooth, q.dev_F_smooth_sum = r["F_smooth"], r["F_smooth_sum"]

    # ---- kappa / c_p / entropy / phase number from file (C:199-292; K:703-919)
    def _entr(self, q, temp, press, tab, n, log_t):
        return O.scalar_interpol(temp, q.dev_entr_temp, press, q.dev_entr_press, tab, int(q.entr_npress),
                                 int(q.entr_ntemp), int(n), log_t=log_t)

    def interpolate_kappa_and_cp(self, q):
        if not isinstance(q.input_kappa_value, str):
            return
        q.dev_kappa_lay = self._entr(q, q.dev_T_lay, q.dev_p_lay, q.dev_entr_kappa, q.nlayer, False)
        q.dev_c_p_lay = self._entr(q, q.dev_T_lay, q.dev_p_lay, q.dev_entr_c_p, q.nlayer, True)
        if q.iso == 0:
            q.dev_kappa_int = self._entr(q, q.dev_T_int, q.dev_p_int, q.dev_entr_kappa, q.ninterface, False)

    def interpolate_entropy(self, q):
        if isinstance(q.input_kappa_value, str):
            q.dev_entropy_lay = self._entr(q, q.dev_T_lay, q.dev_p_lay, q.dev_entr_entropy, q.nlayer, True)

    def interpolate_phase_state(self, q):
        if q.input_kappa_value == "water_atmo":
            q.dev_phase_number_lay = self._entr(q, q.dev_T_lay, q.dev_p_lay, q.dev_entr_phase_number, q.nlayer, False)

    # ---- post-processing
    def integrate_optdepth_transmission(self, q):
        nb, nl, ny = int(q.nbin), int(q.nlayer), int(q.ny)
        if q.iso == 1:
            tb, db = O.integrate_optdepth_transmission_iso(q.dev_trans_wg, q.dev_delta_tau_wg, q.dev_gauss_weight, nb, nl, ny)
        else:
            tb, db, dtc = O.integrate_optdepth_transmission_noniso(
                q.dev_trans_wg_upper, q.dev_trans_wg_lower, q.dev_delta_tau_wg_upper, q.dev_delta_tau_wg_lower,
                q.dev_gauss_weight, q.dev_delta_tau_all_clouds_upper, q.dev_delta_tau_all_clouds_lower, nb, nl, ny)
            q.dev_delta_tau_all_clouds = dtc
        q.dev_trans_band, q.dev_delta_tau_band = tb, db

    def calculate_contribution_function(self, q):
        nb, nl, ny = int(q.nbin), int(q.nlayer), int(q.ny)
        if q.iso == 1:
            tw, cf = O.calc_contr_func(q.dev_trans_wg, None, q.dev_trans_weight_band, q.dev_gauss_weight,
                                       q.dev_planckband_lay, q.epsi, nb, nl, ny)
        else:
            tw, cf = O.calc_contr_func(q.dev_trans_wg_upper, q.dev_trans_wg_lower, q.dev_trans_weight_band,
                                       q.dev_gauss_weight, q.dev_planckband_lay, q.epsi, nb, nl, ny)
        q.dev_trans_weight_band, q.dev_contr_func_band = tw, cf

    def calculate_mean_opacities(self, q):
        r = O.calc_mean_opacities(q.dev_opac_wg_lay, q.dev_abs_cross_all_clouds_lay, q.dev_meanmolmass_lay,
                                  q.dev_planckband_lay, q.dev_opac_interwave, q.dev_opac_deltawave, q.dev_T_lay,
                                  q.dev_gauss_weight, q.dev_gauss_y, int(q.nlayer), int(q.nbin), int(q.ny), q.T_star)
        q.dev_planck_opac_T_pl, q.dev_ross_opac_T_pl = r["planck_T_pl"], r["ross_T_pl"]
        q.dev_planck_opac_T_star, q.dev_ross_opac_T_star = r["planck_T_star"], r["ross_T_star"]
        q.dev_opac_band_lay = r["opac_band"]

    def integrate_beamflux(self, q):
        q.dev_F_dir_tot = O.integrate_beamflux(q.dev_F_dir_band, q.dev_opac_deltawave, int(q.nbin), int(q.ninterface))


def oracle_radiation_loop(m, hsfunc, max_iter=100000, verbose=False):
    """computation.py:851-984 on a mirror (premixed opacities, no plotting / coupling): the NumPy
    end-to-end reference used for T-P parity and as the CPU baseline."""
    oc = OracleCompute()
    m.iter_value = 0
    full = int(m.nlayer) + 1
    while True:
        it = int(m.iter_value)
        oc.interpolate_temperatures(m)
        oc.interpolate_planck(m)
        if it % 10 == 0:
            oc.interpolate_opacities_and_scattering_cross_sections(m)
            oc.interpolate_meanmolmass(m)
            if m.clouds == 1:
                oc.calc_total_g_0_of_gas_and_clouds(m)
            m.dev_scat_trigger = np.zeros_like(m.dev_scat_trigger)
            oc.calculate_transmission(m)
            oc.calculate_delta_z(m)
            m.delta_z_lay = m.dev_delta_z_lay
            m.z_lay = np.zeros(int(m.nlayer))
            m.p_lay = m.dev_p_lay
            hsfunc.calculate_height_z(m)
            m.dev_z_lay = m.z_lay
            oc.calculate_direct_beamflux(m)
        if m.flux_calc_method == "iteration":
            oc.populate_spectral_flux_iteratively(m)
        else:
            oc.solve_for_spectral_fluxes_via_matrix(m)
        oc.integrate_flux(m)
        if m.singlewalk == 1:
            break
        oc.rad_temp_iteration(m)
        done = int(m.dev_abort.sum())
        if verbose and it % 100 == 0:
            print("oracle iter", it, "converged", done, "/", full)
        m.iter_value = it + 1
        if m.iter_value in (m.crit_relaxation_numbers or []):
            hsfunc.relax_radiative_convergence_criterion(m)
        if done == full or m.iter_value > max_iter:
            break
    return m


# ---- on-the-fly mixing launch sites (computation.py:1298-1501), added to OracleCompute
def _interpolate_species_opac(self, q):
    k = O.opac_species_interpol(q.dev_T_lay, q.dev_ktemp, q.dev_p_lay, q.dev_kpress, q.dev_opacity_spec_pretab,
                                int(q.npress), int(q.ntemp), int(q.ny), int(q.nbin), int(q.nlayer))
    q.dev_opac_spec_wg_lay = self._pad(k, q.dev_opac_spec_wg_lay.size)
    if q.iso == 0:
        q.dev_opac_spec_wg_int = O.opac_species_interpol(q.dev_T_int, q.dev_ktemp, q.dev_p_int, q.dev_kpress,
                                                         q.dev_opacity_spec_pretab, int(q.npress), int(q.ntemp),
                                                         int(q.ny), int(q.nbin), int(q.ninterface))


def _add_to_mixed_opacity(self, q, mass_spec, s):
    from helios_b200 import host
    mass = np.float64(mass_spec * host.AMU)
    ro = 0 if (q.kcoeff_mixing == "correlated-k" or "CIA" in q.species_list[s].name) else 1
    nl, ni, nb, ny = int(q.nlayer), int(q.ninterface), int(q.nbin), int(q.ny)
    k = O.add_to_mixed_opac(q.dev_vmr_spec_lay, q.dev_opac_spec_wg_lay, q.dev_opac_wg_lay, q.dev_meanmolmass_lay,
                            q.dev_gauss_weight, q.dev_gauss_y, mass, s, ro, ny, nb, nl)
    q.dev_opac_wg_lay = self._pad(k, q.dev_opac_wg_lay.size)
    if q.iso == 0:
        q.dev_opac_wg_int = O.add_to_mixed_opac(q.dev_vmr_spec_int, q.dev_opac_spec_wg_int, q.dev_opac_wg_int,
                                                q.dev_meanmolmass_int, q.dev_gauss_weight, q.dev_gauss_y, mass, s, ro,
                                                ny, nb, ni)


def _calculate_H2O_Rayleigh_scattering(self, q, s):
    from helios_b200 import host
    mass = np.float64(q.species_list[s].weight * host.AMU)
    q.dev_scat_cross_spec_lay = O.calc_h2o_scat(q.dev_T_lay, q.dev_p_lay, q.dev_opac_wave, q.dev_vmr_spec_lay, mass,
                                                int(q.nbin), int(q.nlayer))
    if q.iso == 0:
        q.dev_scat_cross_spec_int = O.calc_h2o_scat(q.dev_T_int, q.dev_p_int, q.dev_opac_wave, q.dev_vmr_spec_int,
                                                    mass, int(q.nbin), int(q.ninterface))


def _add_to_mixed_scat_cross_sect(self, q):
    q.dev_scat_cross_lay = O.add_to_mixed_scat(q.dev_vmr_spec_lay, q.dev_scat_cross_spec_lay, q.dev_scat_cross_lay,
                                               int(q.nbin), int(q.nlayer))
    if q.iso == 0:
        q.dev_scat_cross_int = O.add_to_mixed_scat(q.dev_vmr_spec_int, q.dev_scat_cross_spec_int,
                                                   q.dev_scat_cross_int, int(q.nbin), int(q.ninterface))


OracleCompute.interpolate_species_opac = _interpolate_species_opac
OracleCompute.add_to_mixed_opacity = _add_to_mixed_opacity
OracleCompute.calculate_H2O_Rayleigh_scattering = _calculate_H2O_Rayleigh_scattering
OracleCompute.add_to_mixed_scat_cross_sect = _add_to_mixed_scat_cross_sect
