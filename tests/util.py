"""Shared helpers of the parity tests."""
import numpy as np

from oracle.pipeline import HostMirror

# Parity bar (BASELINE.json north_star): per-kernel relative error <= 1e-10.
# "Relative" is taken against max(|ref|, 1e-6 * max|ref|) so that entries which are pure cancellation
# residue (e.g. 1 - T^2 in optically thin cells, fluxes that underflow towards 0) are measured against the
# scale of their array instead of against themselves.
RTOL = 1e-10
FLOOR = 1e-6


def rel_err(got, ref):
    got = np.asarray(got, np.float64).reshape(-1)
    ref = np.asarray(ref, np.float64).reshape(-1)
    assert got.size == ref.size, (got.size, ref.size)
    if ref.size == 0:
        return 0.0
    both_nan = np.isnan(got) & np.isnan(ref)
    same_inf = np.isinf(ref) & (got == ref)
    fin = np.isfinite(ref) & ~both_nan
    scale = np.max(np.abs(ref[fin])) if fin.any() else 0.0
    denom = np.maximum(np.abs(ref), FLOOR * scale)
    denom = np.where(denom == 0, 1.0, denom)
    with np.errstate(invalid="ignore"):
        err = np.abs(got - ref) / denom
    err = np.where(both_nan | same_inf, 0.0, err)
    err = np.where(np.isnan(err), np.inf, err)
    return float(err.max())


class Failures(list):
    """collects tolerance violations so that one test run reports every offending stage"""

    def check(self):
        assert not self, "parity violations:\n  " + "\n  ".join(self)


def assert_close(got, ref, name, rtol=RTOL, soft=None):
    e = rel_err(got, ref)
    if soft is not None:
        if not e <= rtol:
            soft.append("%s: relative error %.3e exceeds %.1e" % (name, e, rtol))
        return e
    assert e <= rtol, "%s: relative error %.3e exceeds %.1e" % (name, e, rtol)
    return e


# Arrays whose defining formula is ill-conditioned, so that two correct fp64 evaluations (libdevice with
# FMA contraction vs NumPy without) legitimately differ by more than 1e-10.  They are held to 1e-10
# against the reference's own kernels (same compiler, same libdevice) and to the looser bound below
# against NumPy:
#   G_plus/G_minus: denominator E/eps^2 (E-w0)(1-w0 g0) - 1/mu*^2 cancels to O(w0) (K:168)
#   N = zeta+ zeta- (1 - T^2): cancels to O(delta_tau) in optically thin cells (K:1095)
NUMPY_RTOL = {"N_term": 1e-8, "N_upper": 1e-8, "N_lower": 1e-8}


def g_conditioning(q, name):
    """condition number 1/mu*^2 / |denominator| of the G+/- denominator (K:168, 202) for every cell of the
    G array `name`, from the w0 / g0 arrays on the device"""
    from oracle import helios_oracle as O
    nl, nb, ny = int(q.nlayer), int(q.nbin), int(q.ny)
    sfx = "_upper" if name.endswith("_upper") else "_lower" if name.endswith("_lower") else ""
    w0 = getattr(q, "dev_w_0" + sfx).get()[:nl * nb * ny].reshape(nl, nb, ny)
    if q.clouds == 1:
        gl = q.dev_g_0_tot_lay.get()[:nl * nb].reshape(nl, nb, 1)
        if sfx:
            gi = q.dev_g_0_tot_int.get()[:(nl + 1) * nb].reshape(nl + 1, nb, 1)
            g0 = (gl + gi[1:]) / 2.0 if sfx == "_upper" else (gi[:-1] + gl) / 2.0
        else:
            g0 = gl
    else:
        g0 = float(q.g_0)
    E = O.E_parameter(w0, g0 * np.ones_like(w0), q.i2s_transition) if q.scat_corr == 1 else 1.0
    with np.errstate(divide="ignore", invalid="ignore"):
        denom = E * q.epsi ** -2.0 * (E - w0) * (1.0 - w0 * g0) - q.mu_star ** -2.0
        cond = (q.mu_star ** -2.0) / np.abs(denom)
    return np.where(np.isfinite(cond), cond, np.inf).reshape(-1)


def assert_close_G(q, got, ref, name, rtol=RTOL, soft=None):
    """G+/-: the denominator E/eps^2 (E-w0)(1-w0 g0) - 1/mu*^2 cancels to O(w0) -- exactly, for the default
    eps = 1/2 with a 60 degree beam -- so every correct fp64 evaluation carries a relative error of a few
    ulp times the condition number of that subtraction.  Cells are held to rtol + 64 ulp * cond; cells
    where the subtraction keeps less than ~8 digits (cond > 1e8: with the denominator within a few ulp of
    zero the result is the +-1e8 limiter or rounding noise in the reference itself, and the condition
    estimate is no longer meaningful) are not compared."""
    got = np.asarray(got, np.float64).reshape(-1)
    ref = np.asarray(ref, np.float64).reshape(-1)
    cond = g_conditioning(q, name)
    n = min(got.size, ref.size, cond.size)
    cond, got, ref = cond[:n], got[:n], ref[:n]
    ok = cond <= 1e8
    allowed = rtol + 64 * 2.2e-16 * cond
    scale = np.max(np.abs(ref[ok])) if ok.any() else 0.0
    with np.errstate(invalid="ignore", divide="ignore"):
        err = np.abs(got - ref) / np.maximum(np.abs(ref), FLOOR * scale)
    excess = np.where(ok, err / allowed, 0.0)
    e = float(np.nanmax(excess)) if excess.size else 0.0
    if not e <= 1.0:
        msg = "%s: error %.2fx the conditioning-aware bound (%d of %d cells compared)" % (name, e, int(ok.sum()), n)
        if soft is None:
            raise AssertionError(msg)
        soft.append(msg)
    return float(np.nanmax(np.where(ok & (cond < 1e3), err, 0.0))) if n else 0.0


def dev_names(q):
    return [k for k, v in vars(q).items() if k.startswith("dev_") and hasattr(v, "ptr") and hasattr(v, "get")]


def restore(q, mirror):
    """write a HostMirror snapshot back into the device buffers of q"""
    for k in dev_names(q):
        h = getattr(mirror, k, None)
        d = getattr(q, k)
        if isinstance(h, np.ndarray) and h.size == d.size:
            d.set(h)


def stage_vs_oracle(q, comp, oc, method, outputs, rtol=RTOL, args=(), soft=None):
    """run one Compute method on the GPU and on a NumPy mirror of the same inputs; compare outputs"""
    q.ctx.synchronize()
    m = HostMirror(q)
    getattr(comp, method)(q, *args)
    getattr(oc, method)(m, *args)
    errs = {}
    for name in outputs:
        got = getattr(q, "dev_" + name).get()
        ref = np.asarray(getattr(m, "dev_" + name))
        n = min(got.size, ref.size)
        if isinstance(rtol, dict) and name in rtol and rtol[name] is None:
            continue
        if name.startswith("G_"):
            errs[name] = assert_close_G(q, got, ref, name, RTOL if isinstance(rtol, dict) else rtol, soft)
            continue
        tol = rtol.get(name, RTOL) if isinstance(rtol, dict) else rtol
        errs[name] = assert_close(got.reshape(-1)[:n], ref.reshape(-1)[:n], method + ":" + name,
                                  max(tol, NUMPY_RTOL.get(name, 0.0)), soft)
    return errs


def stage_vs_ref(q, comp, ref, method, outputs, rtol=RTOL, ref_method=None, args=(), soft=None):
    """run the reference kernel(s) and ours from identical device state; compare outputs"""
    q.ctx.synchronize()
    before = HostMirror(q)
    getattr(ref, ref_method or method)(q, *args)
    want = {name: getattr(q, "dev_" + name).get() for name in outputs}
    restore(q, before)
    q.ctx.synchronize()
    getattr(comp, method)(q, *args)
    q.ctx.synchronize()
    errs = {}
    for name in outputs:
        if name.startswith("G_"):
            errs[name] = assert_close_G(q, getattr(q, "dev_" + name).get(), want[name], name + " (vs kernels.cu)", RTOL, soft)
            continue
        tol = rtol.get(name, RTOL) if isinstance(rtol, dict) else rtol
        errs[name] = assert_close(getattr(q, "dev_" + name).get(), want[name], method + ":" + name + " (vs kernels.cu)", tol, soft)
    return errs
