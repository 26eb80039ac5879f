"""GPU: end-to-end parity of the converged state.  The same host loop (Compute.radiation_loop /
convection_loop) is driven once through the product's C-ABI kernels and once through the reference's own
kernels.cu (RefBacked below re-points every launch site at the verbatim cubin).  BASELINE.json's bar: converged
T-P profile within 0.01 K, TOA emission spectrum within 1e-8 relative -- asserted without any allowance, on the
deterministic form of the reference trajectory.  (The reference's own LOOPS, with its nondeterministic atomics, are
run by tests/test_gpu_refloop.py.)"""
import numpy as np
import pytest

from helios_b200 import synthetic
from helios_b200.computation import Compute
from oracle import ref_gpu

pytestmark = pytest.mark.gpu

SMALL = dict(nbin=37, nlayer=24, ntemp=12, npress=8, plancktable_dim=700, plancktable_step=10)

_SITES = ["construct_planck_table", "correct_incident_energy", "interpolate_temperatures", "interpolate_planck",
          "interpolate_opacities_and_scattering_cross_sections", "interpolate_meanmolmass",
          "calc_total_g_0_of_gas_and_clouds", "calculate_transmission", "calculate_delta_z",
          "calculate_direct_beamflux", "populate_spectral_flux_iteratively", "solve_for_spectral_fluxes_via_matrix",
          "integrate_flux", "rad_temp_iteration", "conv_temp_iteration", "interpolate_kappa_and_cp"]


class RefBacked(Compute):
    """Compute's loops with every kernel launch site served by the reference's kernels.cu"""

    def __init__(self, ctx, deterministic_integration=False):
        """deterministic_integration: keep the product's fixed-order band / wavelength sums for `integrate_flux` only.
        The reference's integrate_flux_double adds with CAS atomics in whatever order the hardware schedules (K:2474), so
        its trajectory is not reproducible run to run; with this one launch site swapped the reference-kernel trajectory
        is deterministic and can be compared at the north-star bars without any allowance for run-to-run spread."""
        super().__init__(ctx, verbose=False)
        self._ref = ref_gpu.RefCompute(ctx.device)
        for name in _SITES:
            if deterministic_integration and name == "integrate_flux":
                continue
            setattr(self, name, self._bind(name))

    def prepare_iteration(self, q):  # Compute fuses the two launch sites; the reference has them separately
        self.interpolate_temperatures(q)
        self.interpolate_planck(q)

    def _bind(self, name):
        def call(q, *a):
            self.ctx.synchronize()  # the reference launches on the NULL stream
            getattr(self._ref, name)(q, *a)
        return call

    def _layers_converged(self, q):
        return int(q.dev_abort.get().sum())  # C:927-932


def _run(ctx, config, backed, perturb=0.0, T_intern=None, det=False, limit=None):
    q = synthetic.make_store(config, ctx=ctx, **SMALL)
    if config == "C2":
        q.mu_star = np.float64(np.cos((180 - 50.0) * np.pi / 180.0))
    if T_intern is not None:
        from helios_b200 import host
        q.T_intern = np.float64(T_intern)
        host.calc_F_intern(q)
    if perturb:
        q.T_lay = np.asarray(q.T_lay, np.float64) + perturb
    if limit is not None:
        q.rad_convergence_limit = np.float64(limit)
    synthetic.upload(q)
    comp = RefBacked(ctx, deterministic_integration=det) if backed else Compute(ctx, verbose=False)
    comp.construct_planck_table(q)
    comp.correct_incident_energy(q)
    comp.radiation_loop(q, None, None, None)
    rad_iters = int(q.iter_value)
    T_rad = q.dev_T_lay.get().copy()
    comp.convection_loop(q, None, None, None)
    nl, nb = int(q.nlayer), int(q.nbin)
    return dict(T_rad=T_rad, T=q.dev_T_lay.get().copy(), rad_iters=rad_iters, conv_iters=int(q.iter_value),
                toa=q.dev_F_up_band.get()[nl * nb:(nl + 1) * nb].copy(), F_net=q.dev_F_net.get().copy(),
                F_intern=float(q.F_intern), conv=int(q.convection),
                Fdn_top=float(q.dev_F_down_tot.get()[nl]), limit=float(q.rad_convergence_limit))


def _lockstep(ctx, config, steps=25, det=False, limit=None, threshold=1e-9):
    """both backends from the same start, iteration by iteration: before the adaptive step-size logic
    (comparisons, K:2717-2724) can branch, the temperatures agree to rounding.  Returns (largest relative difference
    over the steps, first iteration at which it exceeded `threshold` or None)."""
    worst, first = 0.0, None
    stores = []
    for backed in (False, True):
        q = synthetic.make_store(config, ctx=ctx, **SMALL)
        if config == "C2":
            q.mu_star = np.float64(np.cos((180 - 50.0) * np.pi / 180.0))
        if limit is not None:
            q.rad_convergence_limit = np.float64(limit)
        synthetic.upload(q)
        comp = RefBacked(ctx, deterministic_integration=det) if backed else Compute(ctx, verbose=False)
        comp.construct_planck_table(q)
        comp.correct_incident_energy(q)
        stores.append((q, comp))
    for it in range(steps):
        Ts = []
        for q, comp in stores:
            q.iter_value = np.int32(it)
            comp.interpolate_temperatures(q)
            comp.interpolate_planck(q)
            if it % 10 == 0:
                comp._refresh_atmosphere(q)
            comp._flux_solve(q)
            comp.rad_temp_iteration(q)
            Ts.append(q.dev_T_lay.get())
        rel = float(np.max(np.abs(Ts[0] - Ts[1]) / Ts[1]))
        if first is None and rel > threshold:
            first = it
        worst = max(worst, rel)
    return worst, first


@pytest.mark.parametrize("config", ["C1", "C2"])
def test_converged_state_at_the_north_star_bars(ctx, config):
    """BASELINE.json's bars with no allowance: converged T-P within 0.01 K, TOA spectrum within 1e-8 relative.
    The reference's kernels behind the same loop, with ONE launch site made deterministic (the band / wavelength sums:
    the reference's CAS-atomic order is the only source of its run-to-run spread), against the product's kernels.

    C1: the two trajectories stay together -- same iteration count (1302) and 1e-7 K with the committed build.
    C2: the pseudo-time controller compares temperatures of consecutive iterations (K:2717-2724); the <= 1e-12
    differences of the kernels flip one of those comparisons after a few hundred iterations (the iteration is printed),
    after which the two runs wander through the convergence basin on different paths.  Both still end INSIDE the basin
    |dF| / F < rad_convergence_limit, and the distance between two points of the basin scales with that limit: at the
    default 1e-8 the deep, optically thick layers (dF/dT ~ 1e-4 of the thin value) are pinned to ~0.02 K only -- the
    reference against itself shows the same (tests/test_gpu_refloop.py).  The bars are therefore asserted on the fixed
    point the criterion approximates, i.e. with the criterion tightened to 1e-11 for both runs (at 1e-10 the two stopping
    points were measured 1.2e-4 K / 1.9e-9 apart with one build and 4.9e-3 K / 2.2e-8 with another: the paths are
    chaotic, the distance scales with the criterion)."""
    if not ref_gpu.available():
        pytest.skip("reference cubin not built")
    lock, _ = _lockstep(ctx, config)  # fully reference-backed (atomics included): 25 iterations agree to rounding
    assert lock < 1e-9, lock
    limit = None if config == "C1" else 1e-11
    if config == "C2":
        _, first = _lockstep(ctx, config, steps=1500, det=True, threshold=1e-7)
        at_default = (_run(ctx, config, backed=False), _run(ctx, config, backed=True, det=True))
        print("\n[rce-det] C2 at the default criterion 1e-8: trajectories separate (rel. dT > 1e-7) at iteration %s; "
              "%d / %d iterations, max |dT| %.2e K" % (first, at_default[0]["rad_iters"], at_default[1]["rad_iters"],
                                                     float(np.max(np.abs(at_default[0]["T"] - at_default[1]["T"])))))
    ours = _run(ctx, config, backed=False, limit=limit)
    ref = _run(ctx, config, backed=True, det=True, limit=limit)
    dT = float(np.max(np.abs(ours["T"] - ref["T"])))
    dT_rad = float(np.max(np.abs(ours["T_rad"] - ref["T_rad"])))
    spec = float(np.max(np.abs(ours["toa"] - ref["toa"]) / np.maximum(np.abs(ref["toa"]), 1e-6 * np.max(np.abs(ref["toa"])))))
    print("\n[rce-det] %s (criterion %s): radiation loop %d / %d iterations, convection loop %d / %d; max |dT| %.2e K (after "
          "the radiation loop %.2e K); TOA spectrum rel. diff %.2e" % (config, ours["limit"], ours["rad_iters"],
                                                                       ref["rad_iters"], ours["conv_iters"], ref["conv_iters"],
                                                                       dT, dT_rad, spec))
    assert ours["rad_iters"] > 50
    assert dT_rad <= 0.01, dT_rad
    assert dT <= 0.01, dT
    assert spec <= 1e-8, spec
    if config == "C1":  # (identical counts with most builds; a reordered sum is enough to move one of the two by a few %)
        assert abs(ours["rad_iters"] - ref["rad_iters"]) <= 0.05 * ref["rad_iters"], (ours["rad_iters"], ref["rad_iters"])
    # radiative equilibrium: F_net == F_intern at every interface of the radiative zone (K:2751, known-answer iii)
    if ours["conv"] == 0:
        scale = ours["Fdn_top"] + ours["F_intern"]
        assert np.max(np.abs(ours["F_net"] - ours["F_intern"])) / scale < 10 * ours["limit"]


def _convection_body(ctx, backed, n_iter):
    """the first n_iter iterations of the radiative-convective loop from a super-adiabatic start profile"""
    q = synthetic.make_store("C2", ctx=ctx, **SMALL)
    q.mu_star = np.float64(np.cos((180 - 50.0) * np.pi / 180.0))
    p = np.asarray(q.p_lay)
    T = np.maximum(3200.0 * (p / p[0]) ** 0.45, 900.0)   # steeper than the dry adiabat (kappa = 2/7) at depth
    q.T_lay = np.append(T, T[0] * 1.05)
    q.max_nr_iterations = n_iter
    synthetic.upload(q)
    comp = RefBacked(ctx) if backed else Compute(ctx, verbose=False)
    comp.construct_planck_table(q)
    comp.correct_incident_energy(q)
    stopped = False
    try:
        comp.convection_loop(q, None, None, None)
    except SystemExit:  # the loop's own iteration guard (C:1157-1164): used here to stop after n_iter iterations
        stopped = True
    return dict(T=q.dev_T_lay.get().copy(), iters=int(q.iter_value), stopped=stopped,
                conv_layers=int(np.sum(q.conv_layer)), F_net=q.dev_F_net.get().copy())


def test_convection_loop_body_matches_reference_kernels(ctx):
    """C:992-1174 actually iterating: host-side convective adjustment between flux solves, the kappa hand-over,
    mark_convective_layers -> conv_temp_iter (K:2768).  With strong internal heating neither backend converges this
    synthetic case within 15,000 iterations (both abort alike), so the loop BODY is compared instead: the first 40
    iterations from a super-adiabatic start, before the step-size logic can branch, agree to rounding."""
    if not ref_gpu.available():
        pytest.skip("reference cubin not built")
    ours = _convection_body(ctx, False, 40)
    ref = _convection_body(ctx, True, 40)
    rel = float(np.max(np.abs(ours["T"] - ref["T"]) / ref["T"]))
    print("\n[rce] convection loop body: %d / %d iterations, %d / %d convective layers, max rel dT %.2e" %
          (ours["iters"], ref["iters"], ours["conv_layers"], ref["conv_layers"], rel))
    assert ours["stopped"] and ref["stopped"], "the loop was expected to run into the iteration guard"
    assert ours["iters"] == ref["iters"] and ours["iters"] > 40
    assert ours["conv_layers"] == ref["conv_layers"] and ours["conv_layers"] > 0
    assert rel < 1e-8, rel


def test_converged_profiles_cross_evaluate_identically(ctx):
    """cross-evaluation: a fresh flux solve of a converged profile gives the same residual
    max|F_intern - F_net| / F whichever backend evaluates it, for profiles converged by either backend.
    (The residual itself is not ~1e-8: the loop stops on the fluxes of the LAST iteration and then still applies
    that iteration's temperature step, whose size shrinks only like |dF|^0.1, K:2694-2700.)"""
    if not ref_gpu.available():
        pytest.skip("reference cubin not built")
    residual = {}
    for first in (True, False):
        a = _run(ctx, "C1", backed=first)
        for second in (True, False):
            q = synthetic.make_store("C1", ctx=ctx, **SMALL)
            q.T_lay = a["T"].copy()
            synthetic.upload(q)
            comp = RefBacked(ctx) if second else Compute(ctx, verbose=False)
            comp.construct_planck_table(q)
            comp.correct_incident_energy(q)
            q.iter_value = np.int32(0)
            comp.interpolate_temperatures(q)
            comp.interpolate_planck(q)
            comp._refresh_atmosphere(q)
            comp._flux_solve(q)
            F_net, Fdn = q.dev_F_net.get(), q.dev_F_down_tot.get()
            residual[(first, second)] = float(np.max(np.abs(q.F_intern - F_net)) / (Fdn[int(q.nlayer)] + q.F_intern))
    print("\n[rce] residual max|F_intern - F_net|/F of a converged profile (converged with ref kernels?, "
          "evaluated with ref kernels?): " + ", ".join("%s=%.2e" % kv for kv in residual.items()))
    for first in (True, False):
        assert abs(residual[(first, False)] - residual[(first, True)]) <= 1e-10 * max(1.0, residual[(first, True)])
    # and the two converged profiles are equivalent: same residual to well within its own size
    assert abs(residual[(True, True)] - residual[(False, True)]) <= 1e-3 * residual[(True, True)]
