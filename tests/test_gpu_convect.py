"""GPU: the device-side forms of the host functions of the radiative-convective loop (csrc/convect.cu)
against their host forms in helios_b200/host.py -- which are pinned to the reference's own host_functions.py by
tests/golden/host_golden.npz (tests/test_host_golden.py).  The algorithms are the same statement for statement; the only
arithmetic difference is libdevice's pow / log10 against NumPy's (last-bit), hence 1e-12 instead of bit equality."""
import types

import numpy as np
import pytest

from helios_b200 import host, synthetic
from helios_b200.computation import Compute

pytestmark = pytest.mark.gpu
SMALL = dict(nbin=37, nlayer=24, ntemp=12, npress=8, plancktable_dim=700, plancktable_step=10)


def _profile(seed, n=40):
    """a bare quant with a super-adiabatic deep atmosphere, an inversion and noise (as tests/golden/make_host_golden.py)"""
    rng = np.random.default_rng(seed)
    q = types.SimpleNamespace()
    q.nlayer, q.ninterface = n, n + 1
    q.p_boa, q.p_toa, q.g = 1e9, 1e-1, 930.0
    q.fl_prec = np.float64
    q.iter_value = 100 if seed % 2 else 6000
    q.T_star = 6117.0 if seed % 3 else 5.0
    q.input_dampara = "automatic" if seed % 4 else "2.5"
    q.dampara = None
    q.F_intern = 45.9
    q.rad_convergence_limit = 1e-4
    q.p_lay, q.p_int = host.calculate_pressure_levels(q)
    q.p_lay, q.p_int = np.array(q.p_lay), np.array(q.p_int)
    kappa = 2.0 / 7.0
    q.kappa_lay = np.full(n, kappa) * (1 + 0.05 * rng.standard_normal(n))
    q.kappa_int = np.full(n + 1, kappa) * (1 + 0.05 * rng.standard_normal(n + 1))
    q.c_p_lay = np.full(n, 2.9e8) * (1 + 0.1 * rng.random(n))
    q.meanmolmass_lay = np.full(n, 2.3 * 1.66e-24) * (1 + 0.05 * rng.random(n))
    T = 1800.0 * (q.p_lay / q.p_lay[0]) ** (0.35 + 0.1 * rng.random())
    T = np.maximum(T, 600.0) + 30.0 * rng.standard_normal(n)
    T[n // 2: n // 2 + 5] += 150.0
    q.T_lay = np.append(T, T[0] * (1.02 + 0.1 * rng.random()))
    q.conv_layer = np.zeros(n + 1, np.int32)
    q.conv_unstable = np.zeros(n + 1, np.int32)
    q.F_add_heat_sum = np.zeros(n)
    q.F_smooth_sum = 1e-3 * rng.standard_normal(n)
    q.F_up_tot = 1e6 * (1 + 0.2 * rng.random(n + 1))
    q.F_down_tot = q.F_up_tot * (1 + 0.01 * rng.standard_normal(n + 1))
    q.F_net = q.F_up_tot - q.F_down_tot
    return q


@pytest.mark.parametrize("seed", range(8))
def test_convective_adjustment_kernel_matches_host(ctx, seed):
    q = _profile(seed)
    n = q.nlayer
    dev = {k: ctx.to_device(np.ascontiguousarray(getattr(q, k), np.float64)) for k in
           ("T_lay", "p_lay", "p_int", "kappa_lay", "kappa_int", "c_p_lay", "meanmolmass_lay", "F_add_heat_sum",
            "F_smooth_sum", "F_down_tot", "F_up_tot")}
    conv_layer, unstable, status = ctx.zeros(n + 1, np.int32), ctx.zeros(n + 1, np.int32), ctx.zeros(4, np.int32)
    dampara = -1.0 if q.input_dampara == "automatic" else float(q.input_dampara)
    ctx.call("convective_adjustment", dev["T_lay"], dev["p_lay"], dev["p_int"], dev["kappa_lay"], dev["kappa_int"],
             dev["c_p_lay"], dev["meanmolmass_lay"], dev["F_add_heat_sum"], dev["F_smooth_sum"], dev["F_down_tot"],
             dev["F_up_tot"], conv_layer, unstable, status, q.F_intern, q.T_star, dampara, q.iter_value, n)
    host.convective_adjustment(q)
    st = status.get()
    assert st[3] == 0 and st[2] == 1, st  # the start profile IS unstable, and stability was reached
    assert np.array_equal(conv_layer.get(), np.asarray(q.conv_layer, np.int32))
    assert np.array_equal(unstable.get(), np.asarray(q.conv_unstable, np.int32))
    got = dev["T_lay"].get()
    assert np.max(np.abs(got - q.T_lay) / q.T_lay) < 1e-12
    # ... and then the marking + radiative-equilibrium test on the adjusted profile
    marked, status2 = ctx.zeros(n + 1, np.int32), ctx.zeros(4, np.int32)
    Fnet = ctx.to_device(q.F_net)
    ctx.call("convection_marks", dev["T_lay"], dev["p_lay"], dev["p_int"], dev["kappa_lay"], dev["kappa_int"], Fnet,
             dev["F_down_tot"], dev["F_add_heat_sum"], dev["F_smooth_sum"], conv_layer, marked, status2, q.F_intern,
             q.rad_convergence_limit, q.iter_value, n)
    host.mark_convective_layers(q, stitching=1)
    q.iter_value = 0  # (silences the host function's progress print)
    eq = host.check_for_radiative_eq(q)
    s2 = status2.get()
    assert np.array_equal(conv_layer.get(), np.asarray(q.conv_layer, np.int32))
    assert np.array_equal(marked.get(), np.asarray(q.marked_red, np.int32))
    assert s2[0] == int(sum(q.converged)) and s2[1] == (n + 1) - int(sum(q.conv_layer)) and s2[2] == int(sum(q.conv_layer))
    assert (s2[0] == s2[1]) == bool(eq)


def _convection_run(ctx, device, n_iter=40):
    q = synthetic.make_store("C2", ctx=ctx, **SMALL)
    q.mu_star = np.float64(np.cos((180 - 50.0) * np.pi / 180.0))
    p = np.asarray(q.p_lay)
    T = np.maximum(3200.0 * (p / p[0]) ** 0.45, 900.0)   # steeper than the dry adiabat (kappa = 2/7) at depth
    q.T_lay = np.append(T, T[0] * 1.05)
    q.max_nr_iterations = n_iter
    synthetic.upload(q)
    comp = Compute(ctx, verbose=False)
    comp.device_convection = device
    comp.construct_planck_table(q)
    comp.correct_incident_energy(q)
    calls0 = ctx.launch_count()
    try:
        comp.convection_loop(q, None, None, None)
    except SystemExit:
        pass
    return dict(T=q.dev_T_lay.get(), iters=int(q.iter_value), conv=np.asarray(q.conv_layer, np.int32).copy(),
                launches=ctx.launch_count() - calls0)


def test_convection_loop_on_the_device_equals_the_host_path(ctx):
    """C:992-1174 with the convective adjustment, the marking and the equilibrium test on the device against the same loop
    with the host functions (and their transfers): same iteration count, same convective layers, temperatures to rounding"""
    dev, hst = _convection_run(ctx, True), _convection_run(ctx, False)
    rel = float(np.max(np.abs(dev["T"] - hst["T"]) / hst["T"]))
    print("\n[convect] %d / %d iterations, %d convective layers, max rel dT %.2e" % (dev["iters"], hst["iters"], int(dev["conv"].sum()), rel))
    assert dev["iters"] == hst["iters"] and dev["iters"] > 40
    assert np.array_equal(dev["conv"], hst["conv"]) and dev["conv"].sum() > 0
    assert rel < 1e-9, rel


def test_vmr_and_meanmolmass_on_the_device(ctx):
    """H:874-959: per-species VMR profiles from (T, log10 P) tables and the mean molecular mass, device vs host"""
    q = synthetic.make_store("C3", ctx=ctx, kcoeff_mixing="correlated-k", n_species=5, **SMALL)
    q.iso = np.int32(0)
    rng = np.random.default_rng(3)
    nt, npr = int(q.ntemp), int(q.npress)
    for s, sp in enumerate(q.species_list):
        if s % 2 == 0:  # FastChem-tabulated species: a (T, P)-dependent VMR table
            sp.source_for_vmr = "FastChem"
            sp.vmr_pretab = (10.0 ** rng.uniform(-8, -1, (nt, npr))).reshape(-1)
        else:
            sp.source_for_vmr = "constant"
    n = int(q.nlayer)
    q.T_lay = np.concatenate([np.linspace(6500.0, 30.0, n), [2300.0]])  # beyond both ends of the temperature grid
    synthetic.upload(q)
    comp = Compute(ctx, verbose=False)
    comp.interpolate_temperatures(q)
    comp.calculate_vmr_and_meanmolmass_on_device(q)
    got_mu = (q.dev_meanmolmass_lay.get(), q.dev_meanmolmass_int.get())
    got_vmr = [(a.get(), b.get()) for a, b in q._vmr_on_device]
    host.calculate_vmr_for_all_species(q)
    host.calculate_meanmolecularmass(q)
    for s, sp in enumerate(q.species_list):
        assert np.max(np.abs(got_vmr[s][0] - sp.vmr_layer) / sp.vmr_layer) < 1e-12, s
        assert np.max(np.abs(got_vmr[s][1] - sp.vmr_interface) / sp.vmr_interface) < 1e-12, s
    assert np.max(np.abs(got_mu[0] - q.meanmolmass_lay) / q.meanmolmass_lay) < 1e-13
    assert np.max(np.abs(got_mu[1] - q.meanmolmass_int) / q.meanmolmass_int) < 1e-13
