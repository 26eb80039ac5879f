import sys, numpy as np
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
from helios_b200 import synthetic, runtime
from helios_b200.computation import Compute
from oracle import ref_gpu
from oracle.pipeline import HostMirror
import test_gpu_parity as T
from util import restore
ctx = runtime.default_context()
for variant in ("C2", "C2_scorr", "C1_beam_geom"):
    q = T._variant(variant, ctx)
    comp = Compute(ctx, verbose=False)
    steps = ["construct_planck_table", "correct_incident_energy", "interpolate_temperatures", "interpolate_planck",
             "interpolate_opacities_and_scattering_cross_sections", "interpolate_meanmolmass"]
    if q.clouds == 1: steps.append("calc_total_g_0_of_gas_and_clouds")
    steps += ["calculate_transmission", "calculate_direct_beamflux"]
    for m in steps: getattr(comp, m)(q)
    ctx.synchronize()
    before = HostMirror(q)
    ref = ref_gpu.RefCompute(0)
    ref.populate_spectral_flux_iteratively(q)
    R = q.dev_F_down_wg.get().copy()
    out = {}
    for mode in (1, 2):
        restore(q, before); ctx.synchronize()
        ctx.set_fband_mode(mode)
        comp.populate_spectral_flux_iteratively(q)
        out[mode] = q.dev_F_down_wg.get().copy()
    ctx.set_fband_mode(0)
    nint = int(q.ninterface); nc = int(q.nbin)*int(q.ny)
    R = R.reshape(nint, nc); A = out[1].reshape(nint, nc); B = out[2].reshape(nint, nc)
    scale = np.abs(R).max()
    print("==", variant, "scale", scale)
    for nm, X in (("column", A), ("layerpar", B)):
        err = np.abs(X - R) / np.maximum(np.abs(R), 1e-6*scale)
        i, c = np.unravel_index(np.argmax(err), err.shape)
        print(nm, "max err %.3e at interface %d col %d (x=%d,y=%d): ref %.6e got %.6e; pure rel %.3e; bitwise equal frac %.4f" % (
            err.max(), i, c, c//int(q.ny), c%int(q.ny), R[i,c], X[i,c], abs(X[i,c]-R[i,c])/abs(R[i,c]), np.mean(X==R)))
        pure = np.abs(X-R)/np.maximum(np.abs(R),1e-300)
        print("   worst pure relative error %.3e ; per-interface max floor-rel:" % pure.max(), np.array2string(err.max(axis=1), precision=1))
