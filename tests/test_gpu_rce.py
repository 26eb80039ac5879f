"""GPU: end-to-end parity of the converged state.  The same host loop (Compute.radiation_loop /
convection_loop) is driven once through the product's C-ABI kernels and once through the reference's own
kernels.cu (RefBacked below re-points every launch site at the verbatim cubin).  BASELINE.json's bar: converged
T-P profile within 0.01 K, TOA emission spectrum within 1e-8 relative."""
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

    def __init__(self, ctx):
        super().__init__(ctx, verbose=False)
        self._ref = ref_gpu.RefCompute(ctx.device)
        for name in _SITES:
            setattr(self, name, self._bind(name))

    def _bind(self, name):
        def call(q, *a):
            self.ctx.synchronize()  # the reference launches on the NULL stream
            getattr(self._ref, name)(q, *a)
        return call

    def _layers_converged(self, q):
        return int(q.dev_abort.get().sum())  # C:927-932


def _run(ctx, config, backed):
    q = synthetic.make_store(config, ctx=ctx, **SMALL)
    if config == "C2":
        q.mu_star = np.float64(np.cos((180 - 50.0) * np.pi / 180.0))
    synthetic.upload(q)
    comp = RefBacked(ctx) if backed else Compute(ctx, verbose=False)
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


@pytest.mark.parametrize("config", ["C1", "C2"])
def test_converged_profile_and_spectrum_match_reference_kernels(ctx, config):
    if not ref_gpu.available():
        pytest.skip("reference cubin not built")
    ours = _run(ctx, config, backed=False)
    ref = _run(ctx, config, backed=True)
    dT = float(np.max(np.abs(ours["T"] - ref["T"])))
    dT_rad = float(np.max(np.abs(ours["T_rad"] - ref["T_rad"])))
    spec = float(np.max(np.abs(ours["toa"] - ref["toa"]) / np.maximum(np.abs(ref["toa"]), 1e-6 * np.max(np.abs(ref["toa"])))))
    print("\n[rce] %s: radiation loop %d (ours) / %d (kernels.cu) iterations, convection loop %d / %d; "
          "max |dT| = %.2e K after the radiation loop, %.2e K at the end; TOA spectrum rel. diff %.2e" %
          (config, ours["rad_iters"], ref["rad_iters"], ours["conv_iters"], ref["conv_iters"], dT_rad, dT, spec))
    assert ours["rad_iters"] > 50, "the loop did not iterate"
    assert dT_rad <= 0.01, dT_rad
    assert dT <= 0.01, dT
    assert spec <= 1e-8, spec
    # radiative equilibrium: F_net == F_intern at every interface of the radiative zone (K:2751, known-answer iii)
    if ours["conv"] == 0:
        scale = ours["Fdn_top"] + ours["F_intern"]
        assert np.max(np.abs(ours["F_net"] - ours["F_intern"])) / scale < 10 * ours["limit"]
