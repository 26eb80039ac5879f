"""Generates tests/golden/host_golden.npz by running the REFERENCE's own source/host_functions.py
(imported from /root/reference; it cannot travel to the GPU box) on seeded inputs.

The reference module imports pycuda and astropy at module level; neither is installed here, so this script
installs inert stand-ins for exactly those imports (test scaffolding only, nothing from them is executed by the
functions exercised below).  The physical constants the stand-in provides are the ones helios_b200/host.py
uses, so both sides evaluate with identical scalars.

    python tests/golden/make_host_golden.py          # needs /root/reference
"""
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
REF = os.environ.get("HELIOS_REFERENCE", "/root/reference")


def install_stubs():
    from helios_b200 import host as H
    for name in ("pycuda", "pycuda.driver", "pycuda.autoinit", "pycuda.gpuarray", "pycuda.compiler"):
        sys.modules[name] = types.ModuleType(name)
    sys.modules["pycuda.compiler"].SourceModule = object
    sys.modules["pycuda"].driver = sys.modules["pycuda.driver"]
    sys.modules["pycuda"].gpuarray = sys.modules["pycuda.gpuarray"]

    class _C(object):
        def __init__(self, v):
            self.cgs = types.SimpleNamespace(value=v)
            self.esu = types.SimpleNamespace(value=v)

    const = types.ModuleType("astropy.constants")
    vals = dict(c=H.C, k_B=H.K_B, h=H.H, R=H.R_UNIV, N_A=6.022140857e23, sigma_sb=H.SIGMA_SB, au=H.AU, u=H.AMU,
                m_e=9.10938356e-28, e=4.80320425e-10, R_sun=H.R_SUN, M_sun=1.98847e33, R_jup=H.R_JUP,
                M_jup=1.89813e30, R_earth=6.3781e8, M_earth=5.97217e27, G=6.6743e-8, sigma_T=6.6524587158e-25)
    for k, v in vals.items():
        setattr(const, k, _C(v))
    astropy = types.ModuleType("astropy")
    astropy.constants = const
    sys.modules["astropy"] = astropy
    sys.modules["astropy.constants"] = const


def make_quant(seed, nlayer=40, kind="hot_bottom"):
    """a bare object with the fields the convection routines touch"""
    rng = np.random.default_rng(seed)
    q = types.SimpleNamespace()
    q.nlayer = nlayer
    q.ninterface = nlayer + 1
    q.p_boa, q.p_toa, q.g = 1e9, 1e-1, 930.0
    q.fl_prec = np.float64
    q.iter_value = 100 if seed % 2 else 6000
    q.T_star = 6117.0 if seed % 3 else 5.0
    q.input_dampara = "automatic" if seed % 4 else "2.5"
    q.dampara = None
    q.F_intern = 45.9
    q.rad_convergence_limit = 1e-4
    return q, rng


def fill_profile(q, rng, H):
    n = q.nlayer
    q.p_lay, q.p_int = H.calculate_pressure_levels(q)
    q.p_lay = np.array(q.p_lay)
    q.p_int = np.array(q.p_int)
    kappa = 2.0 / 7.0
    q.kappa_lay = np.full(n, kappa) * (1 + 0.05 * rng.standard_normal(n))
    q.kappa_int = np.full(n + 1, kappa) * (1 + 0.05 * rng.standard_normal(n + 1))
    q.c_p_lay = np.full(n, 2.9e8) * (1 + 0.1 * rng.random(n))
    q.meanmolmass_lay = np.full(n, 2.3 * 1.66e-24) * (1 + 0.05 * rng.random(n))
    # super-adiabatic deep atmosphere + an inversion + noise
    T = 1800.0 * (q.p_lay / q.p_lay[0]) ** (0.35 + 0.1 * rng.random())
    T = np.maximum(T, 600.0) + 30.0 * rng.standard_normal(n)
    T[n // 2: n // 2 + 5] += 150.0
    q.T_lay = np.append(T, T[0] * (1.02 + 0.1 * rng.random()))
    q.conv_layer = np.zeros(n + 1, np.int32)
    q.conv_unstable = np.zeros(n + 1, np.int32)
    q.F_add_heat_sum = np.zeros(n)
    q.F_smooth_sum = np.zeros(n)
    q.F_down_tot = 1e8 * (0.2 + rng.random(n + 1))
    q.F_up_tot = q.F_down_tot * (1 + 0.02 * rng.standard_normal(n + 1))
    q.F_net = q.F_up_tot - q.F_down_tot
    q.marked_red = np.zeros(n + 1, np.int32)
    q.delta_z_lay = 1e6 * (1 + rng.random(n))
    q.z_lay = np.zeros(n)
    q.planet_type = "gas"


def main():
    install_stubs()
    sys.path.insert(0, REF)
    from source import host_functions as R  # the reference module, unmodified
    out = {}
    for seed in range(8):
        q, rng = make_quant(seed)
        fill_profile(q, rng, R)
        for k in ("p_lay", "p_int", "kappa_lay", "kappa_int", "c_p_lay", "meanmolmass_lay", "T_lay", "F_down_tot",
                  "F_up_tot", "F_net", "delta_z_lay"):
            out["in%d_%s" % (seed, k)] = np.array(getattr(q, k), np.float64)
        out["in%d_scalars" % seed] = np.array([q.iter_value, q.T_star, 0 if q.input_dampara == "automatic" else float(q.input_dampara)])
        R.conv_check(q)
        out["out%d_conv_unstable0" % seed] = q.conv_unstable.copy()
        R.mark_convective_layers(q, stitching=0)
        out["out%d_conv_layer0" % seed] = q.conv_layer.copy()
        R.convective_adjustment(q)
        out["out%d_T_adjusted" % seed] = np.array(q.T_lay, np.float64)
        out["out%d_conv_layer" % seed] = q.conv_layer.copy()
        out["out%d_eq" % seed] = np.array([R.check_for_radiative_eq(q)])
        out["out%d_marked_red" % seed] = q.marked_red.copy()
        out["out%d_converged" % seed] = q.converged.copy()
        R.calculate_height_z(q)
        out["out%d_z_lay" % seed] = q.z_lay.copy()
        R.calculate_conv_flux(q)
        out["out%d_F_net_conv" % seed] = q.F_net_conv.copy()
    q, _ = make_quant(99, nlayer=100)
    q.delta_colmass, q.delta_col_upper, q.delta_col_lower = [], [], []
    R.construct_grid(q)
    for k in ("p_lay", "p_int", "delta_colmass", "delta_col_upper", "delta_col_lower"):
        out["grid_" + k] = np.array(getattr(q, k), np.float64)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "host_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, len(out), "arrays")


if __name__ == "__main__":
    main()
