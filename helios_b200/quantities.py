"""Host-side mirror of the reference's `quantities.Store` (source/quantities.py:29-665).

Same public surface -- `dimensions`, `create_zero_arrays`, `convert_input_list_to_array`,
`copy_host_to_device`, `allocate_on_device`, `copy_device_to_host` and the plain attributes the
unchanged reader / host functions / writers touch -- but every `dev_*` handle is a
`helios_b200.backend.DeviceArray` owned by libhelios_b200.so instead of a PyCUDA gpuarray or raw
`mem_alloc`.  The three tables below replace the reference's ~200 hand-written assignment lines.
"""
import numpy as np

from . import backend

# scalar attributes that the reader / host functions fill (Q:37-129), with the reference defaults
_SCALARS = dict(
    iso=None, nlayer=None, ninterface=None, p_toa=None, p_boa=None, singlewalk=None, scat=None,
    diffusivity=None, convection=None, epsi=None, epsi2=None, f_factor=None, T_intern=None, ntemp=None,
    npress=None, entr_ntemp=None, entr_npress=None, g_0=None, planet=None, g=None, a=None, R_planet=None,
    R_star=None, T_star=None, T_eff_final=None, model=None, real_star=np.int32(0), name=None, foreplay=None,
    realtime_plot=None, prec=None, fl_prec=None, nr_bytes=None, iter_value=None, ny=None, nbin=None,
    nlayer_nbin=None, nlayer_plus2_nbin=None, ninterface_nbin=None, nlayer_wg_nbin=None,
    ninterface_wg_nbin=None, nplanck_grid=None, dir_beam=None, dir_angle=None, mu_star=None, w_0_limit=None,
    w_0_scat_limit=None, delta_tau_limit=None, rad_convergence_limit=None, global_limit=None, n_plot=None,
    energy_correction=None, star_corr_factor=np.int32(1), input_dampara=None, dampara=None, F_intern=None,
    adapt_interval=None, smooth=None, geom_zenith_corr=None, scat_corr=None, input_kappa_value=None,
    approx_f=None, tau_lw=1, planet_type=None, F_sens=0, debug=None, kappa_file_format=np.int32(0),
    i2s_transition=None, flux_calc_method=None, relaxed_criterion_trigger=0, clouds=None, add_heating=None,
    add_heating_path=None, add_heating_file_header_lines=None, add_heating_file_press_name=None,
    add_heating_file_press_unit=None, add_heating_file_data_name=None,
    add_heating_file_data_conv_factor=None, no_atmo_mode=np.int32(0), physical_tstep=None,
    runtime_limit=None, force_start_tp_from_file=None, plancktable_dim=None, plancktable_step=None,
    kcoeff_mixing=None, opacity_mixing=None, coupling=None, coupling_full_output=None,
    coupling_speed_up=None, coupling_iter_nr=None, coupl_tp_write_interval=None,
    coupl_convergence_limit=None, max_nr_iterations=None, wg_nbin=None,
)

# host input arrays copied to the device as they are (Q:467-495, 540-549); "list" ones start as []
_INPUTS = [
    "p_lay", "p_int", "delta_colmass", "delta_col_upper", "delta_col_lower", "ktemp", "kpress", "entr_temp",
    "entr_press", "opac_k", "gauss_y", "gauss_weight", "opac_wave", "opac_deltawave", "opac_interwave",
    "opac_scat_cross", "opac_meanmass", "entr_kappa", "entr_c_p", "entr_phase_number", "entr_entropy",
    "c_p_lay", "kappa_lay", "starflux", "T_lay", "surf_albedo", "abs_cross_all_clouds_lay",
    "scat_cross_all_clouds_lay", "g_0_all_clouds_lay",
]
_INPUTS_NONISO = ["abs_cross_all_clouds_int", "scat_cross_all_clouds_int", "g_0_all_clouds_int", "kappa_int"]
_LIST_INIT = {"delta_colmass", "delta_col_upper", "delta_col_lower", "entr_temp", "entr_press", "entr_kappa",
              "entr_c_p", "entr_phase_number", "entr_entropy", "T_lay"}
# the arrays convert_input_list_to_array() casts (Q:369-398)
_CONVERT = [
    "p_lay", "p_int", "delta_colmass", "delta_col_upper", "delta_col_lower", "ktemp", "kpress", "entr_temp",
    "entr_press", "opac_k", "gauss_y", "gauss_weight", "opac_wave", "opac_deltawave", "opac_interwave",
    "opac_scat_cross", "opac_meanmass", "entr_kappa", "entr_c_p", "entr_entropy", "entr_phase_number",
    "starflux", "T_lay", "surf_albedo", "abs_cross_all_clouds_lay", "abs_cross_all_clouds_int",
    "scat_cross_all_clouds_lay", "scat_cross_all_clouds_int", "g_0_all_clouds_lay", "g_0_all_clouds_int",
]

# zero-initialised arrays that exist on both sides: name -> (size attribute or lambda, dtype, noniso only)
# (Q:414-461 for the host side, Q:499-545 for the device side)
_ZEROS = [
    ("F_up_band", "ninterface_nbin", "f", False), ("F_down_band", "ninterface_nbin", "f", False),
    ("F_dir_band", "ninterface_nbin", "f", False), ("F_up_wg", "ninterface_wg_nbin", "f", False),
    ("F_down_wg", "ninterface_wg_nbin", "f", False), ("F_dir_wg", "ninterface_wg_nbin", "f", False),
    ("Fc_up_wg", "ninterface_wg_nbin", "f", True), ("Fc_down_wg", "ninterface_wg_nbin", "f", True),
    ("Fc_dir_wg", "ninterface_wg_nbin", "f", True), ("F_up_tot", "ninterface", "f", False),
    ("F_down_tot", "ninterface", "f", False), ("F_dir_tot", "ninterface", "f", False),
    ("opac_band_lay", "nlayer_nbin", "f", False), ("opac_wg_lay", "nlayer_wg_nbin", "f", False),
    ("opac_wg_int", "ninterface_wg_nbin", "f", False), ("scat_cross_lay", "nlayer_nbin", "f", False),
    ("scat_cross_int", "ninterface_nbin", "f", False), ("F_net", "ninterface", "f", False),
    ("F_net_diff", "nlayer", "f", False), ("meanmolmass_lay", "nlayer", "f", False),
    ("meanmolmass_int", "ninterface", "f", False), ("planckband_lay", "nlayer_plus2_nbin", "f", False),
    ("planckband_int", "ninterface_nbin", "f", True), ("planck_opac_T_pl", "nlayer", "f", False),
    ("ross_opac_T_pl", "nlayer", "f", False), ("planck_opac_T_star", "nlayer", "f", False),
    ("ross_opac_T_star", "nlayer", "f", False), ("trans_band", "nlayer_nbin", "f", False),
    ("delta_tau_band", "nlayer_nbin", "f", False), ("abort", "nlayer_plus1", "i", False),
    ("entropy_lay", "nlayer", "f", False), ("phase_number_lay", "nlayer", "f", False),
    ("trans_weight_band", "nlayer_nbin", "f", False), ("contr_func_band", "nlayer_nbin", "f", False),
    ("g_0_tot_lay", "nlayer_nbin", "f", False), ("g_0_tot_int", "ninterface_nbin", "f", True),
    ("delta_z_lay", "nlayer", "f", False), ("z_lay", "nlayer", "f", False), ("T_int", "ninterface", "f", False),
    ("scat_trigger", "wg_nbin", "i", False), ("delta_tau_all_clouds", "nlayer_nbin", "f", False),
    ("F_add_heat_lay", "nlayer", "f", False), ("F_add_heat_sum", "nlayer", "f", False),
    ("F_smooth", "nlayer", "f", False), ("F_smooth_sum", "nlayer", "f", False),
]
# device-only work arrays (Q:613-665): name -> (size key, condition)
_DEVICE_ONLY = [
    ("delta_t_prefactor", "nlayer_plus1", "always"), ("T_store", "nlayer_plus1", "always"),
    ("planckband_grid", "nplanck_grid", "always"), ("opac_wg_lay", "nlayer_wg_nbin", "always"),
    ("delta_tau_wg", "nlayer_wg_nbin", "always"), ("trans_wg", "nlayer_wg_nbin", "always"),
    ("w_0", "nlayer_wg_nbin", "always"), ("M_term", "nlayer_wg_nbin", "always"),
    ("N_term", "nlayer_wg_nbin", "always"), ("P_term", "nlayer_wg_nbin", "always"),
    ("G_plus", "nlayer_wg_nbin", "always"), ("G_minus", "nlayer_wg_nbin", "always"),
    ("opac_spec_wg_lay", "nlayer_wg_nbin", "otf"),
    ("delta_tau_wg_upper", "nlayer_wg_nbin", "noniso"), ("delta_tau_wg_lower", "nlayer_wg_nbin", "noniso"),
    ("trans_wg_upper", "nlayer_wg_nbin", "noniso"), ("trans_wg_lower", "nlayer_wg_nbin", "noniso"),
    ("M_upper", "nlayer_wg_nbin", "noniso"), ("N_upper", "nlayer_wg_nbin", "noniso"),
    ("P_upper", "nlayer_wg_nbin", "noniso"), ("M_lower", "nlayer_wg_nbin", "noniso"),
    ("N_lower", "nlayer_wg_nbin", "noniso"), ("P_lower", "nlayer_wg_nbin", "noniso"),
    ("w_0_upper", "nlayer_wg_nbin", "noniso"), ("w_0_lower", "nlayer_wg_nbin", "noniso"),
    ("G_plus_upper", "nlayer_wg_nbin", "noniso"), ("G_plus_lower", "nlayer_wg_nbin", "noniso"),
    ("G_minus_upper", "nlayer_wg_nbin", "noniso"), ("G_minus_lower", "nlayer_wg_nbin", "noniso"),
    ("delta_tau_all_clouds_upper", "nlayer_nbin", "noniso"),
    ("delta_tau_all_clouds_lower", "nlayer_nbin", "noniso"),
    ("opac_spec_wg_int", "ninterface_wg_nbin", "noniso+otf"),
    ("alpha", "matrix_half", "matrix"), ("beta", "matrix_half", "matrix"),
    ("source_term_down", "matrix_half", "matrix"), ("source_term_up", "matrix_half", "matrix"),
    ("c_prime", "matrix_full", "matrix"), ("d_prime", "matrix_full", "matrix"),
]

# what copy_device_to_host() brings back (Q:554-591)
_D2H = [
    "delta_colmass", "F_up_band", "F_down_band", "F_dir_band", "F_up_tot", "F_down_tot", "F_dir_tot",
    "opac_band_lay", "scat_cross_lay", "F_net", "F_net_diff", "p_lay", "p_int", "T_lay", "planckband_lay",
    "planck_opac_T_pl", "ross_opac_T_pl", "planck_opac_T_star", "ross_opac_T_star", "trans_band",
    "delta_tau_band", "meanmolmass_lay", "c_p_lay", "kappa_lay", "entropy_lay", "phase_number_lay",
    "trans_weight_band", "contr_func_band", "g_0_tot_lay", "delta_z_lay", "z_lay", "delta_tau_all_clouds",
    "F_add_heat_sum", "F_smooth_sum",
]
_D2H_NONISO = ["planckband_int", "kappa_int"]


class Store(object):
    """stores parameters, quantities and arrays; owns the device buffers of one atmosphere"""

    def __init__(self, ctx=None):
        self._ctx = ctx
        for k, v in _SCALARS.items():
            setattr(self, k, v)
        # CPU-only bookkeeping (Q:133-143)
        self.T_restart = []
        self.conv_unstable = None
        self.F_net_conv = []
        self.F_ratio = []
        self.marked_red = None
        self.converged = None
        self.add_heat_dens = None
        self.f_all_clouds_lay = None
        self.f_all_clouds_int = None
        self.species_list = []
        self.crit_relaxation_numbers = None
        self.conv_layer = None
        self.kappa_int = None
        for name in _INPUTS + _INPUTS_NONISO:
            if not hasattr(self, name):
                setattr(self, name, [] if name in _LIST_INIT else None)
            setattr(self, "dev_" + name, None)
        for name, _, _, _ in _ZEROS:
            setattr(self, name, None)
            setattr(self, "dev_" + name, None)
        for name, _, _ in _DEVICE_ONLY:
            setattr(self, "dev_" + name, None)
        for name in ("vmr_spec_lay", "vmr_spec_int", "opacity_spec_pretab", "scat_cross_spec_lay",
                     "scat_cross_spec_int", "conv_layer", "marked_red", "opac_int"):
            setattr(self, "dev_" + name, None)

    # ------------------------------------------------------------------ context
    @property
    def ctx(self):
        if self._ctx is None:
            from . import runtime
            self._ctx = runtime.default_context()
        return self._ctx

    # ------------------------------------------------------------------ Q:366-398
    def convert_input_list_to_array(self):
        """converts lists of quantities to arrays"""
        for name in _CONVERT:
            setattr(self, name, np.array(getattr(self, name), self.fl_prec))

    # ------------------------------------------------------------------ Q:400-409
    def dimensions(self):
        """create the correct dimensions of the grid from input parameters"""
        self.nlayer_nbin = np.int32(self.nlayer * self.nbin)
        self.nlayer_plus2_nbin = np.int32((self.nlayer + 2) * self.nbin)
        self.ninterface_nbin = np.int32(self.ninterface * self.nbin)
        self.ninterface_wg_nbin = np.int32(self.ninterface * self.ny * self.nbin)
        # sic: the reference sizes the "nlayer" wg arrays with ninterface as well (Q:407)
        self.nlayer_wg_nbin = np.int32(self.ninterface * self.ny * self.nbin)
        self.wg_nbin = np.int32(self.ny * self.nbin)
        self.nplanck_grid = np.int32((self.plancktable_dim + 1) * self.nbin)

    def _size(self, key):
        if key == "nlayer_plus1":
            return int(self.nlayer) + 1
        if key == "matrix_half":  # Q:655-663
            return int(self.nlayer_wg_nbin) * (1 if self.iso == 1 else 2)
        if key == "matrix_full":  # Q:607-609
            return 2 * int(self.ninterface_wg_nbin) if self.iso == 1 else 4 * int(self.ninterface_wg_nbin) - 2
        return int(getattr(self, key))

    # ------------------------------------------------------------------ Q:411-461
    def create_zero_arrays(self):
        """creates zero arrays of quantities to be used on the GPU with the correct length/dimension"""
        if self.fl_prec is None:
            self.fl_prec = np.float64
        for name, size_key, kind, _ in _ZEROS:
            dt = np.int32 if kind == "i" else self.fl_prec
            setattr(self, name, np.zeros(self._size(size_key), dt))
        self.conv_layer = np.zeros(int(self.nlayer) + 1, np.int32)

    # ------------------------------------------------------------------ Q:463-549
    def copy_host_to_device(self):
        """copies relevant host arrays to device"""
        if np.dtype(self.fl_prec) != np.float64:
            raise ValueError("helios_b200 implements `precision = double` only (kernels.cu:24-32); got %r" % (self.prec,))
        ctx = self.ctx
        names = list(_INPUTS) + (list(_INPUTS_NONISO) if self.iso == 0 else [])
        for name in names:
            host = getattr(self, name)
            setattr(self, "dev_" + name, ctx.to_device(np.asarray(host, dtype=self.fl_prec)))
        for name, _, _, noniso_only in _ZEROS:
            if noniso_only and self.iso != 0:
                continue
            setattr(self, "dev_" + name, ctx.to_device(getattr(self, name)))

    # ------------------------------------------------------------------ Q:593-665
    def allocate_on_device(self):
        """allocate memory for arrays existing only on the GPU"""
        ctx = self.ctx
        otf = self.opacity_mixing == "on-the-fly"
        for name, size_key, cond in _DEVICE_ONLY:
            need = (cond == "always" or (cond == "noniso" and self.iso == 0) or (cond == "otf" and otf)
                    or (cond == "noniso+otf" and self.iso == 0 and otf)
                    or (cond == "matrix" and self.flux_calc_method == "matrix"))
            if need:
                # the reference leaves these uninitialised (cuda.mem_alloc); zero-filling is a superset
                setattr(self, "dev_" + name, ctx.zeros(self._size(size_key), np.float64))

    # ------------------------------------------------------------------ Q:551-591
    def copy_device_to_host(self):
        """copies relevant device arrays to host"""
        for name in _D2H + (_D2H_NONISO if self.iso == 0 else []):
            setattr(self, name, getattr(self, "dev_" + name).get())
