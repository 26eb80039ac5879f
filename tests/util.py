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


def assert_close(got, ref, name, rtol=RTOL):
    e = rel_err(got, ref)
    assert e <= rtol, "%s: relative error %.3e exceeds %.1e" % (name, e, rtol)
    return e


def dev_names(q):
    return [k for k, v in vars(q).items() if k.startswith("dev_") and hasattr(v, "ptr") and hasattr(v, "get")]


def restore(q, mirror):
    """write a HostMirror snapshot back into the device buffers of q"""
    for k in dev_names(q):
        h = getattr(mirror, k, None)
        d = getattr(q, k)
        if isinstance(h, np.ndarray) and h.size == d.size:
            d.set(h)


def stage_vs_oracle(q, comp, oc, method, outputs, rtol=RTOL, args=()):
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
        errs[name] = assert_close(got.reshape(-1)[:n], ref.reshape(-1)[:n], method + ":" + name, rtol)
    return errs


def stage_vs_ref(q, comp, ref, method, outputs, rtol=RTOL, ref_method=None, args=()):
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
        errs[name] = assert_close(getattr(q, "dev_" + name).get(), want[name], method + ":" + name + " (vs kernels.cu)", rtol)
    return errs
