"""GPU parity at the sizes BASELINE.json names -- the shapes bench.py times, not miniatures of them: the persistent-grid
tail of a 7,700-column atmosphere (1,925 CTA tiles on 444 CTA slots), the two-stage wavelength sum of a 1e5-bin
spectrum, size_t offsets of a 128-atmosphere batch (~6 GB per array family), ten species.  Against the reference's own
kernels.cu (cubin), 1e-10.  Opacity tables are kept small (12 x 8 grid points): the table shape is not what these tests
are about, and snapshots of the device state stay cheap."""
import numpy as np
import pytest

from helios_b200 import synthetic, host
from helios_b200.batch import make_batch
from helios_b200.computation import Compute
from oracle import ref_gpu
from util import HostMirror, restore, assert_close, stage_vs_ref, Failures

pytestmark = pytest.mark.gpu
TABLE = dict(ntemp=12, npress=8, plancktable_dim=700, plancktable_step=10)


def _ref(ctx):
    if not ref_gpu.available():
        pytest.skip("reference cubin not built")
    return ref_gpu.RefCompute(ctx.device)


def _profile(n):
    return np.concatenate([2300.0 - 1200.0 * (np.arange(n) / (n - 1.0)) ** 1.5, [2350.0]])


def _setup(q, comp):
    q.iter_value = np.int32(0)
    for m in ["construct_planck_table", "correct_incident_energy", "interpolate_temperatures", "interpolate_planck",
              "interpolate_opacities_and_scattering_cross_sections", "interpolate_meanmolmass"]:
        getattr(comp, m)(q)
    if q.clouds == 1:
        comp.calc_total_g_0_of_gas_and_clouds(q)
    comp.calculate_transmission(q)
    comp.calculate_delta_z(q)
    q.delta_z_lay = q.dev_delta_z_lay.get()
    host.calculate_height_z(q)
    q.dev_z_lay.set(q.z_lay)
    comp.calculate_direct_beamflux(q)


def _integrate_vs_ref(q, comp, ref, bad):
    """band and wavelength sums against the reference's (whose CAS-atomic summation order varies run to run).  F_net is the
    difference of two nearly equal totals: its error is measured against the totals (the scale the code itself uses,
    K:2751), not against itself."""
    stage_vs_ref(q, comp, ref, "integrate_flux", ["F_down_band", "F_up_band", "F_dir_band", "F_down_tot", "F_up_tot"], soft=bad)
    up, dn, net = q.dev_F_up_tot.get(), q.dev_F_down_tot.get(), q.dev_F_net.get()
    assert np.array_equal(net, up - dn)


@pytest.mark.parametrize("config", ["C1", "C2"])
def test_full_size_flux_solve_integration_and_temperature_step(ctx, config):
    """100 layers x 385 bins x 20 Gauss points: the refresh kernels, two consecutive planned flux solves, the band
    integration and the temperature step, each from identical device state against the reference's kernel"""
    ref = _ref(ctx)
    q = synthetic.make_store(config, ctx=ctx, **TABLE)
    if config == "C2":
        q.mu_star = np.float64(np.cos((180 - 50.0) * np.pi / 180.0))
    q.T_lay = _profile(int(q.nlayer))
    synthetic.upload(q)
    comp = Compute(ctx, verbose=False)
    bad = Failures()
    iso = int(q.iso) == 1
    q.iter_value = np.int32(0)
    for m in ["construct_planck_table", "correct_incident_energy", "interpolate_temperatures", "interpolate_planck"]:
        getattr(comp, m)(q)
    stage_vs_ref(q, comp, ref, "interpolate_opacities_and_scattering_cross_sections",
                 ["opac_wg_lay", "scat_cross_lay"] + ([] if iso else ["opac_wg_int", "scat_cross_int"]), soft=bad)
    comp.interpolate_meanmolmass(q)
    if q.clouds == 1:
        comp.calc_total_g_0_of_gas_and_clouds(q)
    stage_vs_ref(q, comp, ref, "calculate_transmission",
                 ["M_term", "N_term", "P_term", "w_0", "delta_tau_wg"] if iso else
                 ["M_upper", "N_lower", "P_upper", "w_0_lower", "delta_tau_wg_upper"], soft=bad)
    comp.calculate_delta_z(q)
    q.delta_z_lay = q.dev_delta_z_lay.get()
    host.calculate_height_z(q)
    q.dev_z_lay.set(q.z_lay)
    stage_vs_ref(q, comp, ref, "calculate_direct_beamflux", ["F_dir_wg"] + ([] if iso else ["Fc_dir_wg"]), soft=bad)
    fl = ["F_down_wg", "F_up_wg"] + ([] if iso else ["Fc_down_wg", "Fc_up_wg"])
    for solve in range(2):
        ctx.synchronize()
        before = HostMirror(q)
        ref.populate_spectral_flux_iteratively(q)
        want = {n: getattr(q, "dev_" + n).get() for n in fl}
        restore(q, before)
        comp.build_flux_plan(q)  # (restore wrote to every buffer: the plan record was dropped)
        assert q._flux_plan_valid
        comp.populate_spectral_flux_iteratively(q)
        for n in fl:
            assert_close(getattr(q, "dev_" + n).get(), want[n], "planned flux solve %d: %s (vs kernels.cu)" % (solve + 1, n), soft=bad)
        q._flux_plan_valid = False
        _integrate_vs_ref(q, comp, ref, bad)
    stage_vs_ref(q, comp, ref, "rad_temp_iteration", ["T_lay", "abort"], soft=bad)
    bad.check()


@pytest.mark.parametrize("mixing", ["RO", "correlated-k"])
def test_full_size_species_mixing(ctx, mixing):
    """C3: ten species at 100 x 385 x 20 -- interpolation, random overlap / correlated-k, scattering cross sections:
    the mixed opacities after the whole species loop against the reference's kernels run over the same loop"""
    ref = _ref(ctx)
    from test_gpu_mixing import _RefSpecies, _species_loop
    q = synthetic.make_store("C3", ctx=ctx, kcoeff_mixing=mixing, **TABLE)
    q.T_lay = _profile(int(q.nlayer))
    synthetic.upload(q)
    comp = Compute(ctx, verbose=False)
    rs = _RefSpecies(ref, q)
    names = ["opac_wg_lay", "scat_cross_lay"]

    def run(engine):
        def call(m, outs, args):
            ctx.synchronize()
            getattr(engine, m)(q, *args)
        _species_loop(q, comp, call)
        ctx.synchronize()
        return {n: getattr(q, "dev_" + n).get() for n in names}

    want = run(rs)
    got = run(comp)
    for n in names:
        assert_close(got[n], want[n], "C3 %s: %s after the species loop (vs kernels.cu)" % (mixing, n))


def _c4(ctx, nbin=100000):
    q = synthetic.make_store("C4", ctx=ctx, nbin=nbin, **TABLE)
    q.T_lay = _profile(int(q.nlayer))
    synthetic.upload(q)
    comp = Compute(ctx, verbose=False)
    _setup(q, comp)
    return q, comp


def _fband_iso(ctx, q, n):
    ctx.call("fband_iso", q.dev_F_down_wg, q.dev_F_up_wg, q.dev_F_dir_wg, q.dev_planckband_lay, q.dev_w_0, q.dev_M_term,
             q.dev_N_term, q.dev_P_term, q.dev_G_plus, q.dev_G_minus, q.dev_surf_albedo, q.dev_g_0_tot_lay, q.g_0, q.singlewalk,
             q.R_star, q.a, q.ninterface, q.nbin, q.f_factor, q.mu_star, q.ny, q.epsi, q.dir_beam, q.clouds, q.scat_corr,
             q.debug, q.i2s_transition, n)


def _fband_iso_planned(ctx, q, n):
    ctx.call("fband_iso_planned", q.dev_F_down_wg, q.dev_F_up_wg, q.dev_fband_plan, q.dev_planckband_lay, q.dev_surf_albedo,
             q.R_star, q.a, q.ninterface, q.nbin, q.f_factor, q.ny, q.dir_beam, n)


def test_full_size_spectrum(ctx):
    """C4: 100 layers x 1e5 bins x 1 point.  1 pass and 11 fused passes (planned sweep) against as many launches of the
    reference's kernel; the band / wavelength sums (~400 x-tiles walked by 11 blocks per interface) against the
    reference's; 1001 fused passes against the column-serial kernel of fband.cu (bit-faithful evaluation order)."""
    ref = _ref(ctx)
    q, comp = _c4(ctx)
    bad = Failures()
    comp.build_flux_plan(q)
    assert q._flux_plan_valid
    ctx.synchronize()
    zero = {n: np.zeros(getattr(q, "dev_" + n).size) for n in ("F_down_wg", "F_up_wg")}
    one = ref.mod.get_function("fband_iso")
    grid = ((int(q.nbin) + 15) // 16, 1, 1)
    i32, f64 = np.int32, np.float64
    for npass in (1, 11):
        for n, z in zero.items():
            getattr(q, "dev_" + n).set(z)
        comp.build_flux_plan(q)
        ctx.synchronize()
        for _ in range(npass):
            one(q.dev_F_down_wg, q.dev_F_up_wg, q.dev_F_dir_wg, q.dev_planckband_lay, q.dev_w_0, q.dev_M_term, q.dev_N_term,
                q.dev_P_term, q.dev_G_plus, q.dev_G_minus, q.dev_surf_albedo, q.dev_g_0_tot_lay, f64(q.g_0), i32(q.singlewalk),
                f64(q.R_star), f64(q.a), i32(q.ninterface), i32(q.nbin), f64(q.f_factor), f64(q.mu_star), i32(q.ny),
                f64(q.epsi), i32(q.dir_beam), i32(q.clouds), i32(q.scat_corr), i32(q.debug), f64(q.i2s_transition),
                block=(16, 16, 1), grid=grid)
        want = {n: getattr(q, "dev_" + n).get() for n in zero}
        for n, z in zero.items():
            getattr(q, "dev_" + n).set(z)
        comp.build_flux_plan(q)
        _fband_iso_planned(ctx, q, npass)
        for n in zero:
            assert_close(getattr(q, "dev_" + n).get(), want[n], "C4 %d pass(es): %s (vs kernels.cu)" % (npass, n), soft=bad)
    _integrate_vs_ref(q, comp, ref, bad)
    # 1001 passes (post-processing with scattering): planned, layer-parallel, against the column-serial kernel
    for n, z in zero.items():
        getattr(q, "dev_" + n).set(z)
    ctx.set_fband_mode(1)
    try:
        _fband_iso(ctx, q, 1001)
    finally:
        ctx.set_fband_mode(0)
    want = {n: getattr(q, "dev_" + n).get() for n in zero}
    for n, z in zero.items():
        getattr(q, "dev_" + n).set(z)
    comp.build_flux_plan(q)
    _fband_iso_planned(ctx, q, 1001)
    for n in zero:
        assert_close(getattr(q, "dev_" + n).get(), want[n], "C4 1001 passes: %s (planned vs column-serial)" % n, soft=bad)
    bad.check()


def test_full_size_batch_of_128(ctx):
    """C5: 128 atmospheres x (100 x 385 x 20) in one launch per kernel: four sampled atmospheres bit for bit against their
    single-atmosphere runs (12 iterations incl. two refreshes), and those singles' flux solve against kernels.cu"""
    ref = _ref(ctx)
    from helios_b200 import sharding
    params = sharding.partition_atmospheres(synthetic.grid_parameters(), 0, 8)[:128]
    stores = synthetic.make_grid_stores(params, config="C1", ctx=ctx, **TABLE)
    n = int(stores[0].nlayer)
    for k, q in enumerate(stores):
        q.T_lay = _profile(n) + 3.0 * (k % 16)
    sample = [0, 37, 90, 127]
    singles = []
    for b in sample:
        q = synthetic.make_store("C1", ctx=ctx, **TABLE, **params[b])
        q.T_lay = _profile(n) + 3.0 * (b % 16)
        singles.append(synthetic.upload(q))
    qb, bcomp = make_batch(stores, ctx)
    del stores
    comp = Compute(ctx, verbose=False)

    def iterate(q, c, batch):
        c.construct_planck_table(q)
        c.correct_incident_energy(q)
        if batch:
            q.enter()
        for it in range(12):
            q.iter_value = np.int32(it)
            c.interpolate_temperatures(q)
            c.interpolate_planck(q)
            if it % 10 == 0:
                c._refresh_atmosphere(q)
            c.populate_spectral_flux_iteratively(q)
            c.integrate_flux(q)
            c.rad_temp_iteration(q)
        if batch:
            q.leave()

    iterate(qb, bcomp, True)
    for b, q in zip(sample, singles):
        iterate(q, comp, False)
        for name in ("F_down_wg", "F_up_wg", "F_up_band", "F_net", "T_lay", "abort", "M_term", "opac_wg_lay"):
            assert np.array_equal(qb.atmosphere(name, b), getattr(q, "dev_" + name).get()), (b, name)
    # ... and a single of the batch against the reference's kernel
    q = singles[1]
    q._flux_plan_valid = False  # (stage_vs_ref rewrites every buffer, which drops the plan record: the unplanned sweep)
    stage_vs_ref(q, comp, ref, "populate_spectral_flux_iteratively", ["F_down_wg", "F_up_wg"])
