"""CPU: closed-form known answers derived from the reference's formulas, for the NumPy oracle."""
import numpy as np
from numpy.polynomial.legendre import leggauss

from helios_b200 import synthetic
from oracle import helios_oracle as O
from oracle.pipeline import OracleCompute, mirror_from_host

TINY = dict(nbin=41, nlayer=12, ntemp=6, npress=5, plancktable_dim=400, plancktable_step=20)


def test_planck_table_integrates_to_stefan_boltzmann():
    """(ii) sum_x dlambda * pi * B_x(T) -> sigma T^4 when the bins cover the spectrum (K:451-453)"""
    edges = np.geomspace(2e-6, 2e-1, 3001)  # 0.02 micron .. 2 mm: covers the Planck curve at these temperatures
    dl = np.diff(edges)
    grid = O.plancktable(edges, dl, 5800.0, 3, 1000)  # rows T = 1, 1001, 2001 and the stellar row
    n = dl.size
    for row, T in ((1, 1001.0), (2, 2001.0), (3, 5800.0)):
        total = float(np.sum(dl * np.pi * grid[row * n:(row + 1) * n]))
        assert abs(total / (O.STEFANBOLTZMANN * T ** 4) - 1.0) < 2e-4, (T, total)


def test_gauss_weights_sum_to_two():
    """(v) sum 1/2 w_y = 1 (H:222)"""
    assert abs(leggauss(20)[1].sum() - 2.0) < 1e-14


def test_pure_absorption_column_closed_form():
    """(i) scat = 0 => w0 = 0, N = 0, M = -1, P = -T => F_down[i] = T F_down[i+1] + pi B (1 - T) at eps = 1/2 (K:1451)"""
    q = synthetic.make_store("C1", ctx=object(), **TINY)
    q.scat = np.int32(0)
    n = int(q.nlayer)
    q.T_lay = np.concatenate([np.linspace(2100.0, 900.0, n), [2200.0]])
    m = mirror_from_host(q)
    oc = OracleCompute()
    m.iter_value = 0
    for step in ("construct_planck_table", "correct_incident_energy", "interpolate_temperatures", "interpolate_planck",
                 "interpolate_opacities_and_scattering_cross_sections", "interpolate_meanmolmass",
                 "calculate_transmission"):
        getattr(oc, step)(m)
    m.dev_z_lay = np.zeros(n)
    oc.calculate_direct_beamflux(m)
    oc.populate_spectral_flux_iteratively(m)
    nb, ny = int(q.nbin), int(q.ny)
    T = np.asarray(m.dev_trans_wg)[:n * nb * ny].reshape(n, nb * ny)
    Fd = np.asarray(m.dev_F_down_wg).reshape(n + 1, nb * ny)
    B = np.repeat(np.asarray(m.dev_planckband_lay).reshape(nb, n + 2), ny, axis=0)
    assert np.all(np.asarray(m.dev_w_0)[:n * nb * ny] == 0.0)
    for i in range(n - 1, -1, -1):
        want = T[i] * Fd[i + 1] + np.pi * B[:, i] * (1.0 - T[i])
        assert np.max(np.abs(Fd[i] - want) / np.maximum(np.abs(want), 1e-30)) < 1e-9, i


def test_random_overlap_with_a_vanishing_species_returns_the_original():
    """(iv) mixing in a species scaled by 0 leaves the k-distribution unchanged (1 % rule -> correlated-k add of zeros,
    K:3297); and the rebinned result of a genuine overlap is sorted and bounded by the extreme sums"""
    gy = 0.5 * leggauss(20)[0] + 0.5
    gw = leggauss(20)[1]
    rng = np.random.default_rng(3)
    base = np.sort(10.0 ** rng.uniform(-4, 1, (16, 20)), axis=1)
    other = np.sort(10.0 ** rng.uniform(-4, 1, (16, 20)), axis=1)
    out = O.add_to_mixed_opac(np.array([0.0]), other.reshape(-1), base.reshape(-1), np.array([1.0]), gw, gy, 1.0, 1, 1,
                              20, 16, 1).reshape(16, 20)
    assert np.array_equal(out, base)
    mixed = O.add_to_mixed_opac(np.array([1.0]), other.reshape(-1), base.reshape(-1), np.array([1.0]), gw, gy, 1.0, 1, 1,
                                20, 16, 1).reshape(16, 20)
    assert np.all(np.diff(mixed, axis=1) >= 0)
    assert np.all(mixed[:, 0] >= base[:, 0] + other[:, 0]) and np.all(mixed[:, -1] <= base[:, -1] + other[:, -1])
