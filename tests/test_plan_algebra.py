"""CPU check of the algebra behind the sweep plan (helios_b200/csrc/fband_plan.cu: plan_half / k_plan_build_noniso).

Between two opacity refreshes only the Planck values of a half-layer change, and the source term of the two-stream
recurrence (K:1640-1691 downward, K:1744-1795 upward) is affine in them:
    S = [2 pi eps (1 - w0) / (E - w0) * planck_term(B_layer, B_interface) + beam] / M = k0 + k1 * B_layer + k2 * B_interface.
This test restates the plan's eight constants per half-layer in NumPy and holds k0 + k1 B_lay + k2 B_int to the
reference's expression (as restated in oracle/helios_oracle.py: fband_noniso) on random coefficients, for both halves,
both directions, the thin-layer (isothermal) fallback and the gradient form.  The CUDA implementation of the same
constants is held to the reference's kernel on the GPU (tests/test_gpu_parity.py)."""
import numpy as np

PI = np.pi


def plan_half(w0, M, N, P, dt, g0, E, Dd, Du, upper, epsi, delta_tau_limit):
    """[a, b, k0d, k1d, k2d, k0u, k1u, k2u] as k_plan_build stores them (k1 multiplies B_layer, k2 B_interface)"""
    invM = 1.0 / M
    fac = 2.0 * PI * epsi * (1.0 - w0) / (E - w0)
    thin = dt < delta_tau_limit
    pre = epsi / (E * (1.0 - w0 * g0))
    cd = (N + (P - M)) * pre / dt
    cu = ((M - P) - N) * pre / dt
    half = 0.5 * ((M + N) - P)
    if upper:  # down: B1 = layer, B2 = interface above; up: B1 = interface above, B2 = layer
        lay_d, int_d, lay_u, int_u = (M + N) + cd, -P - cd, cu - P, (M + N) - cu
    else:      # down: B1 = interface below, B2 = layer; up: B1 = layer, B2 = interface below
        lay_d, int_d, lay_u, int_u = -P - cd, (M + N) + cd, (M + N) - cu, cu - P
    lay_d, int_d, lay_u, int_u = (np.where(thin, half, v) for v in (lay_d, int_d, lay_u, int_u))
    f = invM * fac
    return invM * P, invM * N, invM * Dd, f * lay_d, f * int_d, invM * Du, f * lay_u, f * int_u


def reference_sources(w0, M, N, P, dt, g0, E, Dd, Du, Blay, Bint, upper, epsi, delta_tau_limit):
    """the source terms of one half-layer as the reference writes them (oracle/helios_oracle.py:531-577)"""
    tpe = 2.0 * PI * epsi
    iso_pt = (Bint + Blay) / 2.0 * (N + M - P)
    if upper:   # K:1640-1664 (down), K:1771-1795 (up); the interface is the one above the layer centre
        pgrad = (Blay - Bint) / dt
        non_d = Blay * (M + N) - Bint * P + epsi / (E * (1.0 - w0 * g0)) * (P - M + N) * pgrad
        non_u = Bint * (M + N) - Blay * P + epsi / (E * (1.0 - w0 * g0)) * pgrad * (M - P - N)
    else:       # K:1667-1691 (down), K:1744-1768 (up); the interface is the one below
        pgrad = (Bint - Blay) / dt
        non_d = Bint * (M + N) - Blay * P + epsi / (E * (1.0 - w0 * g0)) * (P - M + N) * pgrad
        non_u = Blay * (M + N) - Bint * P + epsi / (E * (1.0 - w0 * g0)) * pgrad * (M - P - N)
    pt_d = np.where(dt < delta_tau_limit, iso_pt, non_d)
    pt_u = np.where(dt < delta_tau_limit, iso_pt, non_u)
    Sd = 1.0 / M * (tpe * (1.0 - w0) / (E - w0) * pt_d + Dd)
    Su = 1.0 / M * (tpe * (1.0 - w0) / (E - w0) * pt_u + Du)
    return Sd, Su


def _random_cells(rng, n):
    w0 = rng.uniform(0.0, 0.999, n)
    g0 = rng.uniform(-0.2, 0.9, n)
    E = np.where(rng.random(n) < 0.5, 1.0, rng.uniform(1.0, 1.3, n))
    tr = rng.uniform(1e-6, 1.0, n)              # transmission of the half-layer
    zm, zp = rng.uniform(0.0, 0.5, n), rng.uniform(0.5, 1.0, n)
    M = zm * zm * tr * tr - zp * zp             # K:1048-1050
    N = zp * zm * (1.0 - tr * tr)
    P = (zm * zm - zp * zp) * tr
    dt = np.where(rng.random(n) < 0.2, rng.uniform(0.0, 1e-4, n), 10.0 ** rng.uniform(-4, 2, n))
    Dd = -rng.uniform(0.0, 1e3, n) * (rng.random(n) < 0.5)   # beam sources are min(0, .)
    Du = -rng.uniform(0.0, 1e3, n) * (rng.random(n) < 0.5)
    Blay = 10.0 ** rng.uniform(-8, 6, n)
    Bint = Blay * rng.uniform(0.5, 2.0, n)
    return w0, M, N, P, dt, g0, E, Dd, Du, Blay, Bint


def test_source_terms_are_affine_in_the_planck_values():
    rng = np.random.default_rng(20260114)
    epsi, limit = 0.5, 1e-4
    w0, M, N, P, dt, g0, E, Dd, Du, Blay, Bint = _random_cells(rng, 20000)
    for upper in (True, False):
        a, b, k0d, k1d, k2d, k0u, k1u, k2u = plan_half(w0, M, N, P, dt, g0, E, Dd, Du, upper, epsi, limit)
        Sd, Su = reference_sources(w0, M, N, P, dt, g0, E, Dd, Du, Blay, Bint, upper, epsi, limit)
        for S, k0, k1, k2 in ((Sd, k0d, k1d, k2d), (Su, k0u, k1u, k2u)):
            planned = k0 + k1 * Blay + k2 * Bint
            scale = np.abs(k0) + np.abs(k1 * Blay) + np.abs(k2 * Bint)  # the terms may cancel: error relative to them
            assert np.all(np.abs(planned - S) <= 1e-13 * scale + 1e-300)
        # the homogeneous part of the step: F_out = a F_in - b F_opposite + S  is  1/M (P F_in - N F_opposite + ...)
        Fin, Fop = rng.uniform(0, 1e5, w0.size), rng.uniform(0, 1e5, w0.size)
        direct = 1.0 / M * (P * Fin - N * Fop)
        assert np.allclose(a * Fin - b * Fop, direct, rtol=1e-13, atol=0.0)


def test_plan_survives_a_new_temperature_profile():
    """the same constants serve any Planck values: what the 9 iterations between two refreshes rely on"""
    rng = np.random.default_rng(7)
    epsi, limit = 0.5, 1e-4
    w0, M, N, P, dt, g0, E, Dd, Du, Blay, Bint = _random_cells(rng, 2000)
    consts = {u: plan_half(w0, M, N, P, dt, g0, E, Dd, Du, u, epsi, limit) for u in (True, False)}
    for _ in range(3):
        Blay = Blay * rng.uniform(0.3, 3.0, Blay.size)
        Bint = Bint * rng.uniform(0.3, 3.0, Bint.size)
        for upper in (True, False):
            _, _, k0d, k1d, k2d, k0u, k1u, k2u = consts[upper]
            Sd, Su = reference_sources(w0, M, N, P, dt, g0, E, Dd, Du, Blay, Bint, upper, epsi, limit)
            for S, k0, k1, k2 in ((Sd, k0d, k1d, k2d), (Su, k0u, k1u, k2u)):
                scale = np.abs(k0) + np.abs(k1 * Blay) + np.abs(k2 * Bint)
                assert np.all(np.abs(k0 + k1 * Blay + k2 * Bint - S) <= 1e-13 * scale + 1e-300)


def _kogge_stone_from_top(A, B, nch, hoisted):
    """lane j ends with the composition of the affine maps x -> A x + B of lanes j..nch-1 (lane j applied last), as
    scan_from_top / the hoisted form in fband_cp.cu do it; shuffles past the last lane return the lane's own value"""
    L = A.size
    lanes = np.arange(L)
    mA, mB = A.copy(), B.copy()
    if not hoisted:
        d = 1
        while d < L:
            src = np.where(lanes + d < L, lanes + d, lanes)
            oA, oB = mA[src], mB[src]
            ok = lanes + d < nch
            mA, mB = np.where(ok, mA * oA, mA), np.where(ok, mA * oB + mB, mB)
            d *= 2
        return mA, mB
    steps = []      # once per tile: the multiplier each step applies to the incoming B, 0 where the step is off
    d = 1
    while d < L:
        src = np.where(lanes + d < L, lanes + d, lanes)
        ok = lanes + d < nch
        steps.append(np.where(ok, mA, 0.0))
        mA = np.where(ok, mA * mA[src], mA)
        d *= 2
    d, r = 1, 0     # every pass: B parts only, no select
    while d < L:
        src = np.where(lanes + d < L, lanes + d, lanes)
        mB = steps[r] * mB[src] + mB
        d *= 2
        r += 1
    return mA, mB


def test_hoisted_scan_equals_guarded_scan():
    """the pass-invariant half of the Kogge-Stone scans (used for long pass sequences of the isothermal sweep):
    folding the step guards into zero multipliers reproduces the guarded scan exactly, for every chunk count"""
    rng = np.random.default_rng(3)
    for L in (16, 32):
        for nch in range(1, L + 1):
            A = rng.uniform(0.05, 1.0, L)
            B = rng.normal(size=L) * 10.0 ** rng.uniform(-3, 3, L)
            A[nch:], B[nch:] = 1.0, 0.0
            gA, gB = _kogge_stone_from_top(A, B, nch, hoisted=False)
            hA, hB = _kogge_stone_from_top(A, B, nch, hoisted=True)
            assert np.array_equal(gA[:nch], hA[:nch]) and np.array_equal(gB[:nch], hB[:nch])
            # and both are the sequential composition
            for j in range(nch):
                a_, b_ = 1.0, 0.0
                for i in range(nch - 1, j - 1, -1):
                    a_, b_ = A[i] * a_, A[i] * b_ + B[i]
                assert abs(hA[j] - a_) <= 1e-13 * abs(a_) and abs(hB[j] - b_) <= 1e-12 * (abs(b_) + np.abs(B).max())


def test_upward_weights_are_the_downward_ones_swapped_bit_for_bit():
    """fband_plan.cu stores SIX constants per half-layer [a, b, k0d, k0u, k1, k2] instead of eight: the weights of
    (B_layer, B_interface) in the upward source term are the downward ones swapped -- not just mathematically but bit for
    bit in IEEE arithmetic ((M - P) - N == -(N + (P - M)) exactly, and negation commutes with rounding), in the gradient
    form and in the thin-layer fallback, for both halves."""
    rng = np.random.default_rng(20260115)
    w0, M, N, P, dt, g0, E, Dd, Du, _, _ = _random_cells(rng, 50000)
    for upper in (True, False):
        _, _, _, k1d, k2d, _, k1u, k2u = plan_half(w0, M, N, P, dt, g0, E, Dd, Du, upper, 0.5, 1e-4)
        assert np.array_equal(k1u, k2d) and np.array_equal(k2u, k1d)


def test_multiply_high_division_of_the_sweep_launchers():
    """the sweeps map column -> bin (divide by ny) and CTA tile -> atmosphere (divide by the tiles per atmosphere) with
    one multiply-high by floor(2^32 / d) + 1 (fband_plan.cu: bin_of, atm_of, tile_magic); the launchers only allow it
    while n * d < 2^32 for every n that occurs, where it is exact -- checked here at the range ends and on a sweep"""
    for d in (2, 3, 5, 7, 16, 20, 21, 32, 45, 100, 963, 1925, 12500):
        magic = (1 << 32) // d + 1
        limit = ((1 << 32) - 1) // d  # the launchers' bound: n * d < 2^32
        for n in np.unique(np.concatenate([np.arange(0, 4096), np.arange(limit - 4096, limit + 1),
                                           np.random.default_rng(d).integers(0, limit + 1, 20000)])).astype(np.uint64):
            if int(n) * d >= (1 << 32) or n < 0:
                continue
            assert (int(n) * magic) >> 32 == int(n) // d, (d, int(n))


def test_staging_pitch_rule():
    """column pitch of the staged flux rows (CtaShape::cpitch): the smallest pitch >= rs that is (16 / columns) * odd,
    which puts the column starts of a cooperative 8-byte access 16 / columns bank pairs apart (conflict-free for the
    copiers, lanes along the columns, and for the owners, lanes along one column)"""
    def cpitch(rs, ncol):
        st = 1 if ncol >= 16 else 16 // ncol
        v = rs
        while v % (2 * st) != st:
            v += 1
        return v

    for ncol, lanes_per_col in ((4, 8), (8, 4)):  # a warp request: ncol columns x 32 / ncol slots
        for rs in range(2, 34, 2):
            cp = cpitch(rs, ncol)
            assert rs <= cp < rs + 2 * (16 // ncol) and (cp // (16 // ncol)) % 2 == 1
            for l0 in range(0, 32, lanes_per_col):
                offs = [c * cp + l0 + l for l in range(lanes_per_col) for c in range(ncol)]
                for half in (offs[:16], offs[16:]):  # 8-byte accesses are served per half-warp over 16 bank pairs
                    assert len({o % 16 for o in half}) == 16, (ncol, rs, cp)
