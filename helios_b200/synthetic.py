"""Seeded synthetic inputs of HELIOS's table shapes (SURVEY.md 8d / BASELINE.md 3).

The real opacity tables, stellar spectra and Mie files are a Zenodo download and are not available
offline, so every benchmark and parity case runs on inputs generated here.  `make_store(config)`
fills a `quantities.Store` with exactly the attributes the reference's reader (source/read.py) and host
functions would have produced before `keeper.create_zero_arrays()` (helios.py:46-76), for the five
BASELINE.json configurations:

  "C1"  hot Jupiter, premixed corr-k table 385 bins x 20 Gauss points, isothermal layers
  "C2"  same, non-isothermal layers + one cloud deck + non-gray albedo + direct beam at 60 deg
  "C3"  on-the-fly mixing of 10 species (kcoeff_mixing = "RO" or "correlated-k")
  "C4"  opacity-sampling post-processing, 1e5 bins x 1 point
  (C5, the 1024-atmosphere grid, is a set of C1 stores: see `grid_parameters`)

Sizes can be overridden (nbin=, nlayer=, ntemp=, npress=, plancktable_dim=) for fast tests.
"""
import numpy as np
from numpy.polynomial.legendre import leggauss

from . import host
from .quantities import Store

SEED = 20260117


class Species(object):
    """the per-species record of on-the-fly mixing (read.py:1324-1645)"""

    def __init__(self, name, weight, absorbing="yes", scattering="no"):
        self.name = name
        self.weight = weight
        self.absorbing = absorbing
        self.scattering = scattering
        self.source_for_vmr = "constant"
        self.vmr_layer = None
        self.vmr_interface = None
        self.vmr_pretab = None
        self.opacity_pretab = None
        self.scat_cross_sect_layer = None
        self.scat_cross_sect_interface = None


def wavelength_grid(config, nbin=None):
    """bin interfaces in cm"""
    if config == "C4":
        n = int(nbin or 100000)
        return 0.3e-4 * (30.0 / 0.3) ** (np.arange(n + 1) / n)
    if nbin is None or int(nbin) == 385:
        return 0.244e-4 * (51.0 / 50.0) ** np.arange(386)  # R = 50, 0.244 - 500 micron
    n = int(nbin)  # reduced test sizes keep the spectral coverage, not the resolution
    return 0.244e-4 * (51.0 / 50.0) ** (385.0 * np.arange(n + 1) / n)


def k_table(rng, ktemp, kpress, lam_c, gauss_y, spread=True, scale=1.0):
    """monotone-in-y synthetic k-distribution table, flat [t][p][x][y] in cm2/g"""
    nbin = lam_c.size
    a = rng.uniform(-5.0, 1.0, nbin)
    a = np.convolve(np.pad(a, 2, mode="edge"), np.ones(5) / 5.0, mode="valid")
    s = rng.uniform(0.5, 2.5, nbin) if spread else np.zeros(nbin)
    logk = (a[None, None, :, None]
            + 0.5 * np.log10(kpress / 1e6)[None, :, None, None]
            + 1.5 * ((ktemp - 1500.0) / 1500.0)[:, None, None, None]
            + s[None, None, :, None] * (2.0 * gauss_y - 1.0)[None, None, None, :])
    return (scale * 10.0 ** logk).reshape(-1)


def rayleigh_table(ktemp, kpress, lam_c):
    """flat [t][p][x] in cm2"""
    sig = 5e-27 * (lam_c / 1e-4) ** -4.0
    tfac = 1.0 + 0.1 * (ktemp / 1000.0 - 1.0)
    return (tfac[:, None, None] * np.ones(kpress.size)[None, :, None] * sig[None, None, :]).reshape(-1)


def make_store(config="C1", ctx=None, nbin=None, nlayer=100, ntemp=120, npress=28, plancktable_dim=8000,
               plancktable_step=2, kcoeff_mixing="RO", n_species=10, table_scale=1.0, T_star=6117.0, g=930.0,
               T_lay=None, seed=SEED, tables=None, ngauss=20):
    """`tables`: optional dict scale -> (opac_k, opac_scat_cross, opac_meanmass) shared between the stores of a
    grid, filled on first use (the stores then reference ONE host copy per scaling)"""
    rng = np.random.default_rng(seed)
    q = Store(ctx)
    q.name = "synthetic_" + config
    q.prec, q.fl_prec, q.nr_bytes = "double", np.float64, 8
    # ---- run configuration (param.dat defaults, read.py:331-985)
    iterative = config != "C4"
    q.singlewalk = np.int32(0 if iterative else 1)
    q.iso = np.int32(1 if config in ("C1", "C3", "C4") else 0)
    q.energy_correction = np.int32(1 if iterative else 0)
    q.scat = np.int32(1)
    q.diffusivity = 2.0
    q.epsi = np.float64(1.0 / q.diffusivity)
    q.epsi2 = np.float64(0.5)
    q.g_0 = np.float64(0.0)
    q.f_factor = np.float64(0.5)
    q.T_intern = np.float64(30.0)
    q.scat_corr = np.int32(0)
    q.i2s_transition = np.float64(0.1)
    q.debug = np.int32(0)
    q.smooth = np.int32(0)
    q.adapt_interval = np.int32(20)
    q.foreplay = np.int32(0)
    q.physical_tstep = np.float64(0)
    q.runtime_limit = np.float64(86400)
    q.force_start_tp_from_file = 0
    q.rad_convergence_limit = np.float64(1e-8)
    q.crit_relaxation_numbers = [int(1e4), int(2e4)]
    q.max_nr_iterations = 100000
    q.flux_calc_method = "iteration"
    q.realtime_plot = 0
    q.n_plot = 10
    q.coupling = 0
    q.coupl_tp_write_interval = 0
    q.add_heating = 0
    q.planet_type = "gas"
    q.approx_f = 0
    q.input_dampara = "automatic"
    q.no_atmo_mode = np.int32(0)
    q.plancktable_dim = np.int32(plancktable_dim)
    q.plancktable_step = np.int32(plancktable_step)
    q.opacity_mixing = "on-the-fly" if config == "C3" else "premixed"
    q.kcoeff_mixing = kcoeff_mixing
    # direct beam only in C2 (60 deg zenith angle, read.py:897-899)
    q.dir_beam = np.int32(1 if config == "C2" else 0)
    zenith = 60.0
    q.dir_angle = np.float64((180 - zenith) * np.pi / 180.0)
    q.mu_star = np.float64(np.cos(q.dir_angle))
    q.geom_zenith_corr = np.int32(0)
    q.clouds = np.int32(1 if config == "C2" else 0)
    # convective adjustment needs the interface kappa, i.e. non-isothermal layers (computation.py:1004)
    q.convection = np.int32(1 if config == "C2" else 0)
    q.input_kappa_value = np.float64(2.0 / 7.0)
    # ---- planet / star: HD 209458b, blackbody star (planet_database.py:54-61)
    q.planet = "HD_209458b"
    q.R_planet, q.g, q.a, q.R_star, q.T_star = 1.38, float(g), 0.04747, 1.162, float(T_star)
    q.real_star = np.int32(0)
    host.planet_param(q)
    # ---- grid
    q.p_toa, q.p_boa = 1e-1, 1e9
    q.nlayer = np.int32(nlayer)
    q.ninterface = np.int32(nlayer + 1)
    # ---- opacity table axes
    q.ktemp = np.arange(50.0, 50.0 + 50.0 * ntemp, 50.0)[:ntemp] if ntemp == 120 else np.linspace(50.0, 6000.0, ntemp)
    q.kpress = 10.0 ** (np.arange(npress) * (9.0 / (npress - 1)))
    q.ntemp, q.npress = np.int32(ntemp), np.int32(npress)
    edges = wavelength_grid(config, nbin)
    q.opac_interwave = edges
    q.opac_wave = 0.5 * (edges[1:] + edges[:-1])
    q.opac_deltawave = edges[1:] - edges[:-1]
    q.nbin = np.int32(q.opac_wave.size)
    if config == "C4":
        q.gauss_y = np.array([0.0])
    else:
        q.gauss_y = 0.5 * leggauss(int(ngauss))[0] + 0.5
    q.ny = np.int32(q.gauss_y.size)
    host.set_up_numerical_parameters(q)  # gauss_weight and the numerical limits
    shared = tables.get(table_scale) if tables is not None else None
    if shared is not None:
        q.opac_scat_cross, q.opac_meanmass = shared[1], shared[2]
    else:
        q.opac_scat_cross = rayleigh_table(q.ktemp, q.kpress, q.opac_wave)
        q.opac_meanmass = np.full(ntemp * npress, 2.3 * host.AMU)
    if config == "C3":
        q.opac_k = np.zeros(1)
        names = [("H2", 2.01588), ("He", 4.0026), ("H2O", 18.0153), ("CO", 28.01), ("CO2", 44.01), ("CH4", 16.04),
                 ("NH3", 17.031), ("HCN", 27.0253), ("TiO", 63.866), ("VO", 66.9409)][:n_species]
        vmrs = [0.85, 0.14, 1e-3, 5e-4, 1e-4, 1e-4, 1e-5, 1e-5, 1e-6, 1e-6][:n_species]
        for s, ((nm, wt), vmr) in enumerate(zip(names, vmrs)):
            sp = Species(nm, wt, "yes", "yes" if nm in ("H2", "H2O") else "no")
            sp.opacity_pretab = k_table(np.random.default_rng(seed + s), q.ktemp, q.kpress, q.opac_wave, q.gauss_y)
            sp.vmr_layer = np.full(nlayer, vmr)
            sp.vmr_interface = np.full(nlayer + 1, vmr)
            if nm == "H2":
                sig = 5e-27 * (q.opac_wave / 1e-4) ** -4.0
                sp.scat_cross_sect_layer = np.tile(sig, nlayer)
                sp.scat_cross_sect_interface = np.tile(sig, nlayer + 1)
            q.species_list.append(sp)
    elif shared is not None:
        q.opac_k = shared[0]
    else:
        q.opac_k = k_table(rng, q.ktemp, q.kpress, q.opac_wave, q.gauss_y, spread=(config != "C4"), scale=table_scale)
        if tables is not None:
            tables[table_scale] = (q.opac_k, q.opac_scat_cross, q.opac_meanmass)
    # ---- kappa / c_p (read.py:1172-1193) and empty entropy tables
    q.kappa_lay = np.ones(nlayer) * float(q.input_kappa_value) if q.convection == 1 else np.zeros(nlayer)
    q.c_p_lay = np.ones(nlayer) * (host.R_UNIV / float(q.input_kappa_value)) if q.convection == 1 else np.zeros(nlayer)
    q.kappa_int = (np.ones(nlayer + 1) * float(q.input_kappa_value) if q.convection == 1 else np.zeros(nlayer + 1))
    q.entr_temp, q.entr_press, q.entr_kappa, q.entr_c_p, q.entr_entropy, q.entr_phase_number = [], [], [], [], [], []
    q.entr_ntemp = q.entr_npress = np.int32(0)
    # ---- albedo (read.py:1238-1264: scalar values are clamped to [1e-8, 0.999])
    if config == "C2":
        q.surf_albedo = 0.1 + 0.4 * np.exp(-((q.opac_wave - 1e-4) / 5e-5) ** 2)
    else:
        q.surf_albedo = np.ones(int(q.nbin)) * max(1e-8, min(0.999, 0.0))
    q.starflux = np.zeros(int(q.nbin))
    q.dimensions()
    host.construct_grid(q)
    if q.singlewalk == 1:
        # post-processing starts from a given T-P profile (read.py:1274-1322 -> T_restart: surface first, then
        # the layers bottom-up); stand-in: a smooth hot-Jupiter profile
        nl_ = int(q.nlayer)
        q.T_restart = [2350.0] + list(2300.0 - 1200.0 * (np.arange(nl_) / (nl_ - 1.0)) ** 1.5)
    host.initial_temp(q)
    if T_lay is not None:
        q.T_lay = np.array(T_lay, np.float64)
    host.calc_F_intern(q)
    # ---- clouds (clouds.py:179-226 output arrays)
    nl, ni, nb = int(q.nlayer), int(q.ninterface), int(q.nbin)
    if config == "C2":
        def deck(p):
            f_cl = np.exp(-0.5 * (np.log10(np.asarray(p) / 1e5) / 0.5) ** 2)  # log-normal in pressure
            ab = 1e-26 * f_cl[:, None] * np.ones(nb)[None, :]
            sc = 4e-26 * f_cl[:, None] * ((q.opac_wave / 1e-4) ** -1.0)[None, :]
            return ab.reshape(-1), sc.reshape(-1)
        q.abs_cross_all_clouds_lay, q.scat_cross_all_clouds_lay = deck(q.p_lay)
        q.abs_cross_all_clouds_int, q.scat_cross_all_clouds_int = deck(q.p_int)
        q.g_0_all_clouds_lay = np.full(nl * nb, 0.7)
        q.g_0_all_clouds_int = np.full(ni * nb, 0.7)
    else:
        q.abs_cross_all_clouds_lay = np.zeros(nl * nb)
        q.scat_cross_all_clouds_lay = np.zeros(nl * nb)
        q.g_0_all_clouds_lay = np.zeros(nl * nb)
        q.abs_cross_all_clouds_int = np.zeros(ni * nb)
        q.scat_cross_all_clouds_int = np.zeros(ni * nb)
        q.g_0_all_clouds_int = np.zeros(ni * nb)
    return q


def upload(q):
    """helios.py:76-79"""
    q.create_zero_arrays()
    q.convert_input_list_to_array()
    q.copy_host_to_device()
    q.allocate_on_device()
    return q


def make_grid_stores(params, config="C1", ctx=None, **kw):
    """the stores of a grid of atmospheres (C5), sharing one table set per opacity scaling"""
    tables = {}
    return [make_store(config, ctx=ctx, tables=tables, **dict(kw, **p)) for p in params]


def grid_parameters(n_tstar=16, n_logg=16, n_scale=4):
    """C5: T_star 3000-9000 K x log g 2.5-4.0 x table scalings 0.1/1/10/100 -> list of dicts"""
    out = []
    for ts in np.linspace(3000.0, 9000.0, n_tstar):
        for lg in np.linspace(2.5, 4.0, n_logg):
            for sc in (0.1, 1.0, 10.0, 100.0)[:n_scale]:
                out.append(dict(T_star=float(ts), g=float(10 ** lg), table_scale=sc))
    return out
