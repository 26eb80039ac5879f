"""Golden cases shared by the generator (make_ref_kernel_golden.py, runs on the GPU box with the reference's
own kernels.cu) and by the CPU test that pins the NumPy oracle against those vectors (test_oracle_golden.py).

One case = one seeded synthetic Store (tiny sizes) + the launch sequence of source/computation.py
(C:851-984 for the RT iteration, C:1176-1296 for post-processing, C:1454-1501 for on-the-fly mixing).
`run_case` walks that sequence with an *engine* (RefCompute on the device, or OracleCompute on a NumPy mirror)
and calls `after(label, names)` behind every launch site, so that the generator can record the reference's
outputs and the test can compare -- and then overwrite -- the oracle's outputs with them (errors do not
accumulate from stage to stage: every stage starts from the reference's own state).
"""
import numpy as np

from helios_b200 import synthetic, host

TINY = dict(nbin=6, nlayer=9, ntemp=6, npress=5, plancktable_dim=300, plancktable_step=20)

# name -> (config, options)
CASES = {
    "iso": ("C1", {}),
    "iso_scorr_beam": ("C1", dict(scat_corr=1, g_0=0.3, dir_beam=1, geom=1, zenith=80.0)),
    "iso_noscat": ("C1", dict(scat=0)),
    "noniso_cloud_beam": ("C2", dict(zenith=50.0)),
    "noniso_scorr": ("C2", dict(zenith=50.0, scat_corr=1, g_0=0.3)),
    "iso_matrix": ("C1", dict(matrix=1, zenith=50.0, albedo=0.15)),
    "noniso_matrix": ("C2", dict(matrix=1, zenith=50.0, albedo=0.15)),
    "entropy_file": ("C2", dict(zenith=50.0, entropy=1)),
    "mix_ro": ("C3", dict(mixing="RO")),
    "mix_ck_noniso": ("C3", dict(mixing="correlated-k", iso=0)),
    "mix_ro_noniso": ("C3", dict(mixing="RO", iso=0)),
}


def build(case, ctx=None):
    config, opt = CASES[case]
    kw = dict(TINY)
    if config == "C3":
        kw.update(kcoeff_mixing=opt["mixing"], n_species=4)
    q = synthetic.make_store(config, ctx=ctx if ctx is not None else object(), **kw)
    if "iso" in opt:
        q.iso = np.int32(opt["iso"])
    if "zenith" in opt:
        q.mu_star = np.float64(np.cos((180.0 - opt["zenith"]) * np.pi / 180.0))
    if "scat_corr" in opt:
        q.scat_corr = np.int32(1)
        q.g_0 = np.float64(opt["g_0"])
    if "dir_beam" in opt:
        q.dir_beam = np.int32(1)
        q.geom_zenith_corr = np.int32(opt["geom"])
    if "scat" in opt:
        q.scat = np.int32(opt["scat"])
    if "matrix" in opt:
        q.flux_calc_method = "matrix"
    if "albedo" in opt:
        q.surf_albedo = np.ones(int(q.nbin)) * opt["albedo"]
    if "entropy" in opt:
        rng = np.random.default_rng(5)
        nt, npr = 7, 6
        q.input_kappa_value = "water_atmo"
        q.entr_ntemp, q.entr_npress = np.int32(nt), np.int32(npr)
        q.entr_temp = np.linspace(100.0, 4000.0, nt)
        q.entr_press = 10.0 ** np.linspace(-1.0, 9.5, npr)
        q.entr_kappa = rng.uniform(0.1, 0.4, nt * npr)
        q.entr_c_p = rng.uniform(1e7, 4e8, nt * npr)
        q.entr_entropy = rng.uniform(1e8, 1e9, nt * npr)
        q.entr_phase_number = rng.integers(0, 3, nt * npr).astype(np.float64)
    rng = np.random.default_rng(7)
    n = int(q.nlayer)
    q.T_lay = np.concatenate([np.linspace(2300.0, 900.0, n) + rng.uniform(-40, 40, n), [2400.0]])
    return q


ISO_OUT = ["trans_wg", "delta_tau_wg", "M_term", "N_term", "P_term", "G_plus", "G_minus", "w_0",
           "delta_tau_all_clouds", "scat_trigger"]
NONISO_OUT = [b + s for s in ("_upper", "_lower") for b in
              ("trans_wg", "delta_tau_wg", "M", "N", "P", "G_plus", "G_minus", "w_0", "delta_tau_all_clouds")] + ["scat_trigger"]


class DeviceIO(object):
    """state lives in the Store's DeviceArrays"""

    @staticmethod
    def get(q, name):
        return getattr(q, "dev_" + name).get()

    @staticmethod
    def put(q, name, arr):
        arr = np.ascontiguousarray(arr)
        dev = getattr(q, "dev_" + name, None)
        if dev is not None and hasattr(dev, "set") and dev.size == arr.size and dev.dtype == arr.dtype:
            dev.set(arr)
        else:
            setattr(q, "dev_" + name, q.ctx.to_device(arr))
        q.ctx.synchronize()


class MirrorIO(object):
    """state lives in NumPy arrays on a HostMirror"""

    @staticmethod
    def get(m, name):
        return np.asarray(getattr(m, "dev_" + name))

    @staticmethod
    def put(m, name, arr):
        setattr(m, "dev_" + name, np.array(arr))


def run_case(case, q, engine, io, after, ref_signature=False):
    """`ref_signature`: RefCompute.add_to_mixed_opacity takes (mass in grams, s, ro_method)"""
    config, opt = CASES[case]
    iso = int(q.iso) == 1
    q.iter_value = np.int32(0)

    def site(method, names, *args):
        getattr(engine, method)(q, *args)
        after("%s" % method if not args else "%s%s" % (method, args[-1:]), names)

    def height():
        q.delta_z_lay = io.get(q, "delta_z_lay").copy()
        q.z_lay = np.zeros(int(q.nlayer))
        q.p_lay = io.get(q, "p_lay")
        host.calculate_height_z(q)
        io.put(q, "z_lay", np.asarray(q.z_lay, np.float64))

    if config == "C3":
        site("interpolate_temperatures", ["T_int"])
        for kind in ("lay", "int"):  # host_functions.calculate_meanmolecularmass with constant VMRs
            nn = int(q.nlayer) if kind == "lay" else int(q.ninterface)
            mu = np.zeros(nn)
            for sp in q.species_list:
                mu += np.asarray(sp.vmr_layer if kind == "lay" else sp.vmr_interface) * sp.weight * host.AMU
            io.put(q, "meanmolmass_" + kind, mu)
        for s, sp in enumerate(q.species_list):
            io.put(q, "vmr_spec_lay", np.asarray(sp.vmr_layer, np.float64))
            if not iso:
                io.put(q, "vmr_spec_int", np.asarray(sp.vmr_interface, np.float64))
            io.put(q, "opacity_spec_pretab", np.asarray(sp.opacity_pretab, np.float64))
            tag = "[%d]" % s
            getattr(engine, "interpolate_species_opac")(q)
            after("interpolate_species_opac" + tag, ["opac_spec_wg_lay"] + ([] if iso else ["opac_spec_wg_int"]))
            if ref_signature:
                ro = 0 if (q.kcoeff_mixing == "correlated-k" or "CIA" in sp.name) else 1
                engine.add_to_mixed_opacity(q, np.float64(sp.weight * host.AMU), s, ro)
            else:
                engine.add_to_mixed_opacity(q, sp.weight, s)
            after("add_to_mixed_opacity" + tag, ["opac_wg_lay"] + ([] if iso else ["opac_wg_int"]))
            if sp.scattering == "yes":
                if sp.name == "H2O":
                    io.put(q, "scat_cross_spec_lay", np.zeros(int(q.nbin) * int(q.nlayer)))
                    if not iso:
                        io.put(q, "scat_cross_spec_int", np.zeros(int(q.nbin) * int(q.ninterface)))
                    engine.calculate_H2O_Rayleigh_scattering(q, s)
                    after("calculate_H2O_Rayleigh_scattering" + tag,
                          ["scat_cross_spec_lay"] + ([] if iso else ["scat_cross_spec_int"]))
                else:
                    io.put(q, "scat_cross_spec_lay", np.asarray(sp.scat_cross_sect_layer, np.float64))
                    if not iso:
                        io.put(q, "scat_cross_spec_int", np.asarray(sp.scat_cross_sect_interface, np.float64))
                engine.add_to_mixed_scat_cross_sect(q)
                after("add_to_mixed_scat_cross_sect" + tag, ["scat_cross_lay"] + ([] if iso else ["scat_cross_int"]))
        return

    site("construct_planck_table", ["planckband_grid"])
    site("correct_incident_energy", ["planckband_grid"])
    q.iter_value = np.int32(0)
    if "entropy" in opt:
        site("interpolate_temperatures", ["T_int"])
        site("interpolate_kappa_and_cp", ["kappa_lay", "c_p_lay", "kappa_int"])
        site("interpolate_entropy", ["entropy_lay"])
        site("interpolate_phase_state", ["phase_number_lay"])
        return
    matrix = q.flux_calc_method == "matrix"
    for it in range(2):
        q.iter_value = np.int32(it)
        site("interpolate_temperatures", ["T_int"])
        site("interpolate_planck", ["planckband_lay"] + ([] if iso else ["planckband_int"]))
        if it == 0:
            site("interpolate_opacities_and_scattering_cross_sections",
                 ["opac_wg_lay", "scat_cross_lay"] + ([] if iso else ["opac_wg_int", "scat_cross_int"]))
            site("interpolate_meanmolmass", ["meanmolmass_lay"] + ([] if iso else ["meanmolmass_int"]))
            if q.clouds == 1:
                site("calc_total_g_0_of_gas_and_clouds", ["g_0_tot_lay"] + ([] if iso else ["g_0_tot_int"]))
            io.put(q, "scat_trigger", np.zeros(int(q.nbin) * int(q.ny), np.int32))
            site("calculate_transmission", ISO_OUT if iso else NONISO_OUT)
            site("calculate_delta_z", ["delta_z_lay"])
            height()
            site("calculate_direct_beamflux", ["F_dir_wg"] + ([] if iso else ["Fc_dir_wg"]))
        fl = ["F_down_wg", "F_up_wg"] + ([] if iso else ["Fc_down_wg", "Fc_up_wg"])
        if matrix:
            site("solve_for_spectral_fluxes_via_matrix", fl)
        else:
            site("populate_spectral_flux_iteratively", fl)
        site("integrate_flux", ["F_down_band", "F_up_band", "F_dir_band", "F_down_tot", "F_up_tot"])
        site("rad_temp_iteration", ["T_lay", "abort", "T_store", "delta_t_prefactor", "F_net_diff"])
    q.iter_value = np.int32(2)
    red = np.zeros(int(q.nlayer) + 1, np.int32)
    red[3] = 1  # first radiative layer above a convective zone (C:1135-1137)
    io.put(q, "marked_red", red)
    site("conv_temp_iteration", ["T_lay", "T_store", "delta_t_prefactor", "F_net_diff"])
    site("integrate_optdepth_transmission", ["trans_band", "delta_tau_band"] + ([] if iso else ["delta_tau_all_clouds"]))
    site("calculate_contribution_function", ["trans_weight_band", "contr_func_band"])
    site("calculate_mean_opacities", ["planck_opac_T_pl", "ross_opac_T_pl", "planck_opac_T_star", "ross_opac_T_star",
                                      "opac_band_lay"])
    site("integrate_beamflux", ["F_dir_tot"])
