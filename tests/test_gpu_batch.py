"""GPU: batched atmospheres (helios_b200/batch.py, helios_ctx_set_batch) against the same atmospheres run one
by one through the single-atmosphere path.  Every kernel is deterministic and the per-atmosphere arithmetic is
the same code, so the comparison is BIT FOR BIT -- per launch site, and for the converged profiles and
iteration counts of the whole radiation loop."""
import numpy as np
import pytest

from helios_b200 import synthetic
from helios_b200.batch import make_batch
from helios_b200.computation import Compute

pytestmark = pytest.mark.gpu

SMALL = dict(nbin=37, nlayer=24, ntemp=12, npress=8, plancktable_dim=700, plancktable_step=10)
GRID = [dict(T_star=5000.0, g=500.0, table_scale=1.0), dict(T_star=6117.0, g=930.0, table_scale=1.0),
        dict(T_star=8000.0, g=2000.0, table_scale=10.0), dict(T_star=4200.0, g=930.0, table_scale=10.0),
        dict(T_star=6117.0, g=3000.0, table_scale=0.1)]


def _stores(config, ctx, grid=GRID):
    out, tables = [], {}
    for k, p in enumerate(grid):
        q = synthetic.make_store(config, ctx=ctx, **SMALL, **p)
        if config == "C2":
            q.mu_star = np.float64(np.cos((180 - 50.0) * np.pi / 180.0))
        # atmospheres with the same scaling share one table object (-> one copy in HBM)
        key = p["table_scale"]
        if key in tables:
            assert np.array_equal(tables[key], q.opac_k)
            q.opac_k = tables[key]
        else:
            tables[key] = q.opac_k
        n = int(q.nlayer)
        q.T_lay = np.concatenate([np.linspace(2000.0 + 50 * k, 900.0 + 20 * k, n), [2100.0 + 50 * k]])
        out.append(q)
    # the Rayleigh and mean-molecular-mass tables do not depend on the scaling: share them where the k-table is shared
    for q in out:
        first = next(o for o in out if o.opac_k is q.opac_k)
        q.opac_scat_cross, q.opac_meanmass = first.opac_scat_cross, first.opac_meanmass
    return out


def _iterate(q, comp, n_iter, batch):
    comp.construct_planck_table(q)
    comp.correct_incident_energy(q)
    if batch:
        q.enter()
    for it in range(n_iter):
        q.iter_value = np.int32(it)
        comp.interpolate_temperatures(q)
        comp.interpolate_planck(q)
        if it % 10 == 0:
            comp._refresh_atmosphere(q)
        comp.populate_spectral_flux_iteratively(q)
        comp.integrate_flux(q)
        comp.rad_temp_iteration(q)
    if batch:
        q.leave()


CHECK = ["T_int", "planckband_lay", "opac_wg_lay", "scat_cross_lay", "meanmolmass_lay", "delta_z_lay", "z_lay",
         "F_dir_wg", "F_down_wg", "F_up_wg", "F_down_band", "F_up_band", "F_dir_band", "F_down_tot", "F_up_tot", "F_net",
         "F_net_diff", "T_lay", "abort", "T_store", "delta_t_prefactor", "scat_trigger"]
CHECK_ISO = ["trans_wg", "delta_tau_wg", "M_term", "N_term", "P_term", "G_plus", "G_minus", "w_0"]
CHECK_NONISO = ["planckband_int", "opac_wg_int", "scat_cross_int", "meanmolmass_int", "Fc_dir_wg", "Fc_down_wg", "Fc_up_wg",
                "g_0_tot_lay", "g_0_tot_int", "M_upper", "M_lower", "N_upper", "P_lower", "G_plus_upper", "G_minus_lower",
                "w_0_upper", "w_0_lower", "delta_tau_wg_upper", "delta_tau_all_clouds_lower", "trans_wg_lower"]


@pytest.mark.parametrize("config", ["C1", "C2"])
def test_batched_launch_sites_equal_single_runs_bit_for_bit(ctx, config):
    singles = _stores(config, ctx)
    comp = Compute(ctx, verbose=False)
    for q in singles:
        synthetic.upload(q)
        _iterate(q, comp, 12, batch=False)
    qb, bcomp = make_batch(_stores(config, ctx), ctx)
    assert qb.ntables == 3 and list(qb.table_index) == [0, 0, 1, 1, 2]
    _iterate(qb, bcomp, 12, batch=True)
    names = CHECK + (CHECK_ISO if config == "C1" else CHECK_NONISO)
    for b, q in enumerate(singles):
        for name in names:
            want = getattr(q, "dev_" + name).get()
            got = qb.atmosphere(name, b)
            assert got.size == want.size, (name, got.size, want.size)
            same = (got == want) | (np.isnan(got) & np.isnan(want))
            assert same.all(), "atmosphere %d, %s: %d of %d entries differ (max abs diff %.3e)" % (
                b, name, int((~same).sum()), same.size, float(np.nanmax(np.abs(got - want))))


@pytest.mark.parametrize("graph,fused", [(True, True), (True, False), (False, True)])
def test_batched_radiation_loop_reproduces_every_single_run(ctx, graph, fused):
    """the whole loop: device-side iteration counter + convergence latch, blocks of 10 iterations replayed as one
    CUDA graph (graph=True) or launched eagerly (graph=False), against the host-driven single-atmosphere loop"""
    singles = _stores("C1", ctx)
    iters = []
    for q in singles:
        synthetic.upload(q)
        comp = Compute(ctx, verbose=False)
        comp.construct_planck_table(q)
        comp.correct_incident_energy(q)
        comp.radiation_loop(q, None, None, None)
        iters.append(int(q.iter_value))
    qb, bcomp = make_batch(_stores("C1", ctx), ctx)
    bcomp.construct_planck_table(qb)
    bcomp.correct_incident_energy(qb)
    bcomp.radiation_loop(qb, graph=graph, fused=fused)
    print("\n[batch] graph=%s fused=%s: iterations to convergence, single runs %s, batched %s; batched loop %.1f ms for %d "
          "atmospheres (%d iterations run)" % (graph, fused, iters, [int(v) for v in qb.converged_at],
                                               bcomp.stats["radiation_loop_ms"], qb.nbatch, int(qb.iter_value)))
    assert [int(v) for v in qb.converged_at] == iters
    for b, q in enumerate(singles):
        for name in ("T_lay", "F_up_band", "F_down_wg", "F_net", "abort", "T_store", "delta_t_prefactor"):
            assert np.array_equal(qb.atmosphere(name, b), getattr(q, "dev_" + name).get()), (b, name)


def test_single_atmosphere_on_the_device_loop(ctx):
    """nbatch = 1: the device-side loop (graph replay) for ONE atmosphere, non-isothermal layers with clouds and a
    direct beam, against the host-driven loop"""
    grid = [dict(T_star=6117.0, g=930.0, table_scale=1.0)]
    q = _stores("C2", ctx, grid)[0]
    synthetic.upload(q)
    comp = Compute(ctx, verbose=False)
    comp.construct_planck_table(q)
    comp.correct_incident_energy(q)
    comp.radiation_loop(q, None, None, None)
    qb, bcomp = make_batch(_stores("C2", ctx, grid), ctx)
    bcomp.construct_planck_table(qb)
    bcomp.correct_incident_energy(qb)
    bcomp.radiation_loop(qb)
    print("\n[batch] one C2 atmosphere: host loop %d iterations in %.1f ms, device loop %d iterations in %.1f ms" %
          (int(q.iter_value), comp.stats["radiation_loop_ms"], int(qb.converged_at[0]), bcomp.stats["radiation_loop_ms"]))
    assert int(qb.converged_at[0]) == int(q.iter_value)
    for name in ("T_lay", "F_up_band", "Fc_down_wg", "F_net", "abort"):
        assert np.array_equal(qb.atmosphere(name, 0), getattr(q, "dev_" + name).get()), name
    np.testing.assert_array_equal(qb.dev_z_lay.get(), q.dev_z_lay.get())


def test_unbatched_entry_points_refuse_batch_mode(ctx):
    from helios_b200.backend import HeliosError
    qb, bcomp = make_batch(_stores("C1", ctx, GRID[:2]), ctx)
    qb.enter()
    try:
        with pytest.raises(HeliosError):
            bcomp.integrate_beamflux(qb)
        with pytest.raises(HeliosError):  # dimensions must be the declared ones
            ctx.call("temp_inter", qb.dev_T_lay, qb.dev_T_int, int(qb.ninterface) + 1)
    finally:
        qb.leave()


def test_array_released_during_a_capture(ctx):
    """a host garbage collector may release a device array or a graph while launches are being recorded: the free
    (which synchronises) and the graph destruction are deferred to the end of the capture instead of invalidating it"""
    a = ctx.to_device(np.arange(1000.0))
    victim = ctx.to_device(np.zeros(1 << 16))
    b = ctx.zeros(1000)
    with ctx.capture() as g:
        del victim  # -> helios_buf_free inside the capture
        b.copy_from(a)
    g.launch()
    ctx.synchronize()
    assert np.array_equal(b.get(), np.arange(1000.0))
    b.fill_zero()
    with ctx.capture() as g:  # rebinding `g` releases the first graph inside the second capture
        b.copy_from(a)
    g.launch()
    ctx.synchronize()
    assert np.array_equal(b.get(), np.arange(1000.0))
