"""Host-side numerics the RT loops need between kernel launches.

In drop-in mode these come from the reference's unchanged `source/host_functions.py` (H: below); this
module provides the same functions, written from scratch, so that the backend also runs standalone
(bench, tests, batch driver) where the reference tree is absent.  Function names and the fields they
read/write on the `quant` (Store) object are the reference's; behaviour is pinned against the reference
module by tests/golden/host_*.npz (generated with tests/golden/make_host_golden.py).
"""
import math

import numpy as np
from numpy.polynomial.legendre import leggauss

# cgs constants.  The device code carries CODATA-2014 literals (kernels.cu:36-41); the host side of the
# reference takes astropy's.  Standalone we use one consistent set on both sides.
C = 29979245800.0
K_B = 1.38064852e-16
H = 6.62607004e-27
R_UNIV = 8.3144598e7
SIGMA_SB = 5.6703669999999995e-5
AMU = 1.6605390666e-24
AU = 1.495978707e13
R_SUN = 6.957e10
R_JUP = 7.1492e9


# ------------------------------------------------------------------------------ set-up (H:33-48, 164-222, 714-735)
def planet_param(quant, read=None):
    """unit conversion of the planetary / stellar / orbital parameters (H:33-48)"""
    if quant.g < 10:
        quant.g = np.float64(10 ** quant.g)
    quant.a = np.float64(quant.a * AU)
    quant.R_planet = np.float64(quant.R_planet * R_JUP)
    quant.R_star = np.float64(quant.R_star * R_SUN)
    quant.T_star = np.float64(max(quant.T_star, 2.7))


def set_up_numerical_parameters(quant):
    """numerical limits and the Gauss-Legendre weights (H:209-222)"""
    quant.w_0_limit = np.float64(1.0 - 1e-10)
    quant.w_0_scat_limit = np.float64(1e-3)
    quant.delta_tau_limit = np.float64(1e-4)
    quant.gauss_weight = leggauss(int(quant.ny))[1]


def calculate_pressure_levels(quant):
    """log-equidistant staggered grid: even levels are interfaces, odd ones layer centres (H:714-724)"""
    n = int(quant.nlayer)
    ratio = quant.p_toa / quant.p_boa
    levels = [quant.p_boa * ratio ** (k / (2 * n - 1)) for k in range(2 * n)]
    p_layer = levels[1::2]
    p_interface = levels[0::2]
    p_interface.append(quant.p_toa * ratio ** (1 / (2 * n - 1)))
    return p_layer, p_interface


def construct_grid(quant):
    """pressure grid and column masses (H:727-735)"""
    quant.p_lay, quant.p_int = calculate_pressure_levels(quant)
    quant.delta_colmass, quant.delta_col_upper, quant.delta_col_lower = [], [], []
    for i in range(int(quant.nlayer)):
        quant.delta_colmass.append((quant.p_int[i] - quant.p_int[i + 1]) / quant.g)
        quant.delta_col_upper.append((quant.p_lay[i] - quant.p_int[i + 1]) / quant.g)
        quant.delta_col_lower.append((quant.p_int[i] - quant.p_lay[i]) / quant.g)


def model_T_eff(quant):
    geom = (quant.R_star / quant.a) ** 0.5 * quant.T_star
    return (1.0 - quant.dir_beam) * quant.f_factor ** 0.25 * geom + quant.dir_beam * abs(quant.mu_star) ** 0.25 * geom


def initial_temp(quant, read=None):
    """isothermal start at max(T_eff, 500 K) for iterative runs (H:164-184); restart profiles are the
    reader's business and must already sit in quant.T_restart"""
    if quant.singlewalk == 0 and (quant.force_start_tp_from_file == 0 or quant.physical_tstep == 0):
        quant.T_lay = np.ones(int(quant.nlayer) + 1) * max(model_T_eff(quant), 500)
    else:
        quant.T_lay = np.append(quant.T_restart[1:], quant.T_restart[0])


def calc_F_intern(quant):
    quant.F_intern = SIGMA_SB * quant.T_intern ** 4.0  # H:203-206


def relax_radiative_convergence_criterion(quant):
    quant.rad_convergence_limit *= 10.0  # H:243-248
    quant.relaxed_criterion_trigger = 1


# ------------------------------------------------------------------------------ altitude (H:673-698)
def calculate_height_z(quant):
    """altitude of the layer centres: above the 10-bar level for gas planets, above ground otherwise"""
    n = int(quant.nlayer)
    dz = quant.delta_z_lay
    z = quant.z_lay
    if quant.planet_type == "gas":
        ref = max(i for i in range(n) if quant.p_lay[i] >= 1e7)
        z[ref] = 0
        for i in range(ref + 1, n):
            z[i] = z[i - 1] + 0.5 * dz[i - 1] + 0.5 * dz[i]
        for i in range(ref - 1, -1, -1):
            z[i] = z[i + 1] - 0.5 * dz[i + 1] - 0.5 * dz[i]
    elif quant.planet_type in ("rocky", "no_atmosphere"):
        z[0] = 0.5 * dz[0]
        for i in range(1, n):
            z[i] = z[i - 1] + 0.5 * dz[i - 1] + 0.5 * dz[i]


def calc_add_heating_flux(quant):
    """H:701-711"""
    quant.F_add_heat_lay = quant.add_heat_dens * quant.delta_z_lay
    quant.F_add_heat_sum = np.cumsum(quant.F_add_heat_lay)


# ------------------------------------------------------------------------------ convection (H:251-635)
def _adiabat_T_above(quant, i, slack):
    """temperature a parcel from layer i would have at layer i+1 along the dry adiabat, exponents
    stretched by `slack` (H:348-350 / H:557-559)"""
    mid = quant.T_lay[i] * (quant.p_int[i + 1] / quant.p_lay[i]) ** (quant.kappa_lay[i] * slack)
    return mid * (quant.p_lay[i + 1] / quant.p_int[i + 1]) ** (quant.kappa_int[i + 1] * slack)


def _adiabat_T_surface(quant, slack):
    return quant.T_lay[quant.nlayer] * (quant.p_lay[0] / quant.p_int[0]) ** (quant.kappa_int[0] * slack)


def conv_check(quant):
    """flags layer pairs whose lapse rate exceeds the adiabat (H:337-365)"""
    n = int(quant.nlayer)
    flag = np.zeros(n + 1, np.int32)
    for i in range(n - 1):
        if quant.p_lay[i] <= 1e1:  # the uppermost atmosphere is left alone
            break
        if quant.T_lay[i + 1] < _adiabat_T_above(quant, i, 1 + 1e-6):
            flag[i] = flag[i + 1] = 1
    if quant.T_lay[0] < _adiabat_T_surface(quant, 1 + 1e-6):
        flag[n] = flag[0] = 1
    quant.conv_unstable = flag


def mark_convective_layers(quant, stitching):
    """marks where convection dominates (H:545-582); note that layers above the 10 dyn/cm2 cut keep
    their previous marks"""
    n = int(quant.nlayer)
    mark = quant.conv_layer
    mark[n] = 0
    mark[0] = 0
    for i in range(n - 1):
        if quant.p_lay[i] <= 1e1:
            break
        if quant.T_lay[i + 1] < _adiabat_T_above(quant, i, 1 - 1e-6):
            mark[i] = mark[i + 1] = 1
        else:
            mark[i + 1] = 0
    for i in range(n - 1):  # no kink at the top edge of a zone
        if quant.T_lay[i + 1] > quant.T_lay[i]:
            mark[i] = 0
    if quant.T_lay[0] < _adiabat_T_surface(quant, 1 - 1e-6):
        mark[n] = mark[0] = 1
    if stitching == 1 and quant.iter_value > 5000:
        stitching_convective_zone_holes(quant)


def _runs(indices):
    """first / last members of the maximal runs of consecutive integers in a sorted sequence"""
    members = set(int(v) for v in indices)
    starts = [int(v) for v in indices if int(v) - 1 not in members]
    ends = [int(v) for v in indices if int(v) + 1 not in members]
    return starts, ends


def stitching_convective_zone_holes(quant):
    """closes radiative gaps thinner than a scale height between convective zones (H:585-635)"""
    n = int(quant.nlayer)
    mark = quant.conv_layer
    starts, ends = [], []
    for i in range(n):
        if mark[i] != 1:
            continue
        below = mark[i - 1] if i > 0 else mark[n]
        if below == 0:
            starts.append(i)
        if i == n - 1 or mark[i + 1] == 0:
            ends.append(i)
    if mark[n] == 1:
        starts = [-1] + starts
        if mark[0] == 0:
            ends = [-1] + ends
    if len(starts) != len(ends):
        print("Error in stitching calculation. Aborting...")
        raise SystemExit()
    for z in range(len(starts) - 1):
        p_top = quant.p_lay[starts[z + 1]]
        p_bot = quant.p_lay[ends[z]] if ends[z] != -1 else quant.p_int[0]
        if p_top / p_bot > 1 / np.e:
            for m in range(ends[z] + 1, starts[z + 1]):
                mark[m] = 1


def conv_correct(quant, fudging):
    """replaces unstable lapse rates by adiabats that conserve the zone's enthalpy (H:368-506)"""
    n = int(quant.nlayer)
    todo = [i for i in range(n + 1) if quant.conv_unstable[i] == 1 or quant.conv_layer[i] == 1]
    if n in todo:  # the surface sits below layer 0
        todo = [-1] + todo[:-1]
    starts, ends = _runs(todo)
    if len(starts) != len(ends):
        print("Error in convective calculation. Aborting...")
        raise SystemExit()
    nz = len(starts)
    fudge = np.ones(nz)
    if fudging == 1:
        for z in range(nz):
            probe = None
            for m in range(z, nz):
                if m != nz - 1:
                    p_top = quant.p_lay[starts[m + 1]]
                    p_bot = quant.p_lay[ends[m]] if ends[m] != -1 else quant.p_int[0]
                    if p_top / p_bot < 1 / np.e:  # a radiative zone thicker than a scale height follows
                        probe = int((ends[m] + starts[m + 1]) / 2)
                        break
                else:
                    probe = int(0.8 * ends[m] + 0.2 * (quant.ninterface - 1))
            if quant.input_dampara == "automatic":
                if quant.T_star > 10:
                    quant.dampara = 0.5 if z < nz - 1 else 4.0
                else:
                    quant.dampara = 8.0
            else:
                quant.dampara = float(quant.input_dampara)
            f = ((quant.F_intern + quant.F_add_heat_sum[probe - 1] + quant.F_smooth_sum[probe - 1]
                  + quant.F_down_tot[probe]) / quant.F_up_tot[probe]) ** (1.0 / quant.dampara)
            fudge[z] = min(1.01, max(0.99, f))
    for z in range(nz):
        lo, hi = max(0, starts[z]), max(0, ends[z])
        num = 0
        den = 0
        climb = 1  # adiabatic temperature ratio between interface `lo` and interface i
        ratio = {}
        for i in range(lo, hi + 1):
            to_centre = (quant.p_lay[i] / quant.p_int[i]) ** quant.kappa_int[i]
            weight = quant.c_p_lay[i] / quant.meanmolmass_lay[i]
            dp = quant.p_int[i] - quant.p_int[i + 1]
            num += weight * quant.T_lay[i] * dp
            den += climb * (to_centre * quant.c_p_lay[i] / quant.meanmolmass_lay[i] * dp)
            ratio[i] = climb * to_centre
            climb = climb * (to_centre * (quant.p_int[i + 1] / quant.p_lay[i]) ** quant.kappa_lay[i])
        theta = num / den
        theta *= fudge[z]
        for i in range(lo, hi + 1):
            quant.T_lay[i] = theta * ratio[i]
        if starts[z] == -1:
            quant.T_lay[n] = theta


def convective_adjustment(quant):
    """iterate check/correct to stability, then one fudged correction (H:509-538)"""
    conv_check(quant)
    while sum(quant.conv_unstable) > 0:
        mark_convective_layers(quant, stitching=0)
        conv_correct(quant, fudging=0)
        conv_check(quant)
    mark_convective_layers(quant, stitching=1)
    conv_correct(quant, fudging=1)


def check_for_radiative_eq(quant):
    """local radiative-equilibrium test over the radiative layers (H:251-286)"""
    n = int(quant.nlayer)
    quant.converged = np.zeros(n + 1, np.int32)
    quant.marked_red = np.zeros(n + 1, np.int32)
    scale = quant.rad_convergence_limit * (quant.F_down_tot[n] + quant.F_intern)
    for i in range(n + 1):
        if quant.T_lay[i] == 0:
            print("WARNING WARNING WARNING: Found zero temperature at layer:", i, quant.T_lay[i])
        if quant.conv_layer[i] != 0:
            continue
        if i < n:
            miss = abs(quant.F_intern + quant.F_add_heat_sum[i] + quant.F_smooth_sum[i] - quant.F_net[i + 1])
        else:
            miss = abs(quant.F_intern - quant.F_net[0])
        if miss < scale:
            quant.converged[i] = 1
        else:
            quant.marked_red[i] = 1
    n_rad = (n + 1) - sum(quant.conv_layer)
    if quant.iter_value % 100 == 1:
        print("Number of radiative layers converged: {:d} out of {:d}.".format(int(sum(quant.converged)), int(n_rad)))
    return 1 if sum(quant.converged) == n_rad else 0


def give_feedback_on_convergence(quant):
    """prints the energy imbalance in the radiative zones (H:289-318)"""
    n = int(quant.nlayer)
    rad = [i for i in range(n + 1) if quant.conv_layer[i] == 0]
    if n in rad:
        rad = [-1] + rad[:-1]
    starts, ends = _runs(rad)
    norm = quant.F_down_tot[n] + quant.F_intern
    for z in range(len(starts)):
        if z < len(starts) - 1:
            k = int((starts[z] + ends[z] + 1) / 2)
            miss = abs(quant.F_intern + quant.F_add_heat_sum[k - 1] - quant.F_net[k]) / norm
            print("Radiative energy imbalance in intermediate rad. layers is {:.3e} and should be less than {:.1e}".format(miss, quant.rad_convergence_limit))
        else:
            miss = abs(quant.F_intern + quant.F_add_heat_sum[n - 1] - quant.F_net[n]) / norm
            print("Global energy imbalance is {:.3e} and should be less than {:.1e}".format(miss, quant.rad_convergence_limit))


def calculate_conv_flux(quant):
    """convective net flux carried where layers are convective (H:638-651)"""
    quant.F_net_conv = np.zeros(int(quant.ninterface), np.float64)
    for i in range(1, int(quant.ninterface)):
        if quant.conv_layer[i - 1] == 1:
            quant.F_net_conv[i] = quant.F_intern + quant.F_add_heat_sum[i - 1] + quant.F_smooth_sum[i - 1] - quant.F_net[i]
    if quant.conv_layer[quant.nlayer] == 1:
        quant.F_net_conv[0] = quant.F_intern - quant.F_net[0]


def calc_F_ratio(quant):
    """planet-to-star flux ratio per bin (H:654-670)"""
    quant.F_ratio = []
    if quant.T_star > 10:
        geom = (quant.R_planet / quant.R_star) ** 2
        nl, nb = int(quant.nlayer), int(quant.nbin)
        for x in range(nb):
            star = np.pi * quant.planckband_lay[nl + x * (nl + 2)] / quant.star_corr_factor
            quant.F_ratio.append(geom * quant.F_up_band[x + nl * nb] / star if star != 0 else 0)


# ------------------------------------------------------------------------------ on-the-fly mixing (H:874-959, 1050-1056)
def interpolate_grid_to_lay_or_int(log_press, temp, vmr_2D, log_press_profile, temp_profile):
    """bilinear evaluation of a (T, log P) table along a profile (H:904-910: RectBivariateSpline with
    kx = ky = 1 is exactly piecewise-bilinear interpolation, clamped at the grid edges)"""
    temp = np.asarray(temp, np.float64)
    log_press = np.asarray(log_press, np.float64)
    p = np.clip(np.asarray(log_press_profile, np.float64), log_press[0], log_press[-1])
    # the reference walks range(len(log_press_profile)) (H:908): T_lay carries the surface as an extra last entry
    t = np.clip(np.asarray(temp_profile, np.float64)[:p.size], temp[0], temp[-1])
    it = np.clip(np.searchsorted(temp, t, side="right") - 1, 0, temp.size - 2)
    ip = np.clip(np.searchsorted(log_press, p, side="right") - 1, 0, log_press.size - 2)
    ft = (t - temp[it]) / (temp[it + 1] - temp[it])
    fp = (p - log_press[ip]) / (log_press[ip + 1] - log_press[ip])
    v = np.asarray(vmr_2D, np.float64)
    out = (v[it, ip] * (1 - ft) * (1 - fp) + v[it + 1, ip] * ft * (1 - fp) + v[it, ip + 1] * (1 - ft) * fp
           + v[it + 1, ip + 1] * ft * fp)
    return list(out)


def calculate_vmr_for_all_species(quant):
    """vertical VMR profiles of the FastChem-tabulated species (H:874-901)"""
    quant.T_lay = quant.dev_T_lay.get()
    quant.T_int = quant.dev_T_int.get()
    quant.p_lay = quant.dev_p_lay.get()
    quant.p_int = quant.dev_p_int.get()
    log_p_lay, log_p_int, log_kpress = np.log10(quant.p_lay), np.log10(quant.p_int), np.log10(quant.kpress)
    for sp in quant.species_list:
        if sp.source_for_vmr == "FastChem":
            table = sp.vmr_pretab.reshape((int(quant.ntemp), int(quant.npress)))
            sp.vmr_layer = interpolate_grid_to_lay_or_int(log_kpress, quant.ktemp, table, log_p_lay, quant.T_lay)
            if quant.iso == 0:
                sp.vmr_interface = interpolate_grid_to_lay_or_int(log_kpress, quant.ktemp, table, log_p_int, quant.T_int)
            sp.vmr_layer = np.array(sp.vmr_layer, quant.fl_prec)
            sp.vmr_interface = np.array(sp.vmr_interface, quant.fl_prec)


def calc_meanmolmass(quant, type="layer"):
    """VMR-weighted mean molecular mass in grams (H:927-959)"""
    n = int(quant.nlayer) if type == "layer" else int(quant.ninterface)
    total_w = np.zeros(n)
    total_v = np.zeros(n)
    for sp in quant.species_list:
        if "CIA" in sp.name or sp.name in ("H-_ff", "He-"):
            continue
        vmr = np.asarray(sp.vmr_layer if type == "layer" else sp.vmr_interface, np.float64)[:n]
        total_w += vmr * sp.weight
        total_v += vmr
    return np.array(total_w / total_v * AMU, quant.fl_prec)


def calculate_meanmolecularmass(quant):
    """H:913-924"""
    quant.meanmolmass_lay = calc_meanmolmass(quant, "layer")
    quant.dev_meanmolmass_lay = quant.ctx.to_device(quant.meanmolmass_lay)
    if quant.iso == 0:
        quant.meanmolmass_int = calc_meanmolmass(quant, "interface")
        quant.dev_meanmolmass_int = quant.ctx.to_device(quant.meanmolmass_int)


def nullify_opac_scat_arrays(quant):
    """H:1050-1056: the reference re-uploads host zeros; clearing in place needs no PCIe traffic"""
    for name in ("opac_wg_lay", "opac_wg_int", "scat_cross_lay", "scat_cross_int"):
        getattr(quant, "dev_" + name).fill_zero()


def temp_calcs(quant):
    """H:187-200"""
    geom = (quant.R_star / quant.a) ** 0.5 * quant.T_star
    return (0.25 ** 0.25 * geom, 0.667 ** 0.25 * geom, model_T_eff(quant),
            (quant.F_down_tot[quant.ninterface - 1] / SIGMA_SB) ** 0.25,
            (quant.F_up_tot[quant.ninterface - 1] / SIGMA_SB) ** 0.25)
