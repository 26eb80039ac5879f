"""CPU oracle for the HELIOS radiative-transfer hot path -- TEST INFRASTRUCTURE, NOT PRODUCT.

An fp64 NumPy restatement of the reference's device code (`/root/reference/source/kernels.cu`, cited
below as K:line), vectorised over the (wavelength, Gauss-point) columns with Python loops over layers.
Only `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference` legs of
`bench.py` may import this module; nothing under `helios_b200/` does.

Parity status: PINNED against outputs of the reference itself.  The reference ships no tests, golden vectors or
fixtures for this path and its Python side cannot be imported here (PyCUDA/h5py/astropy
absent), so the pin is the reference's own kernels.cu, compiled verbatim to `oracle/_ref/` and launched with the
block/grid shapes of computation.py (oracle/ref_gpu.py):
  * tests/golden/ref_kernels_golden.npz holds its outputs at every launch site for 11 seeded cases (generated on a
    B200 by tests/golden/make_ref_kernel_golden.py); tests/test_oracle_golden.py (CPU) holds this module to them,
    stage by stage, at 1e-10 with four documented conditioning exceptions;
  * tests/test_gpu_parity.py / test_gpu_mixing.py compare it live with the same kernels on the GPU box;
  * tests/test_oracle_kat.py checks closed-form known answers derived from the reference's formulas.

Array conventions are the reference's: "wg" arrays flat [i][x][y] (y fastest, K:1076), band arrays flat
[i][x] (K:2456), Planck arrays flat [x][i] (i fastest, K:940/1004).  Everything is float64.
Arithmetic follows the reference's evaluation order where it matters; NumPy does not contract
multiply-adds, so agreement with device code is to a few ulp, not bitwise.
"""
import numpy as np

# physical constants: the device literals (K:36-41)
PI = 3.141592653589793
HCONST = 6.62607004e-27
CSPEED = 29979245800.0
KBOLTZMANN = 1.38064852e-16
STEFANBOLTZMANN = 5.6703669999999995e-5
AMU = 1.6605390666e-24

f8 = np.float64


# ------------------------------------------------------------------------------- Planck table
def analyt_planck(n, y1, y2):
    """K:95-105"""
    dn = float(n)
    return (np.exp(-dn * y2) * ((y2 * y2 * y2) / dn + 3.0 * (y2 * y2) / (dn * dn) + 6.0 * y2 / (dn * dn * dn)
                                + 6.0 / (dn * dn * dn * dn))
            - np.exp(-dn * y1) * ((y1 * y1 * y1) / dn + 3.0 * (y1 * y1) / (dn * dn) + 6.0 * y1 / (dn * dn * dn)
                                  + 6.0 / (dn * dn * dn * dn)))


def plancktable(lambda_edge, deltalambda, Tstar, dim, step):
    """K:362-416 (the 10 launches of computation.py:46-58 taken together).
    -> planck_grid flat [(dim+1) * nwave], row t<dim at T = t*step+1, row dim at Tstar."""
    lambda_edge = np.asarray(lambda_edge, f8)
    deltalambda = np.asarray(deltalambda, f8)
    nwave = deltalambda.size
    T = np.concatenate([np.arange(dim, dtype=f8) * step + 1.0, [float(Tstar)]])[:, None]  # [row,1]
    kh = KBOLTZMANN / HCONST
    with np.errstate(over="ignore", divide="ignore", invalid="ignore"):
        D = 2.0 * ((kh * kh * kh) * KBOLTZMANN * (T * T * T * T)) / (CSPEED * CSPEED)
        y_top = HCONST * CSPEED / (lambda_edge[None, 1:] * KBOLTZMANN * T)
        y_bot = HCONST * CSPEED / (lambda_edge[None, :-1] * KBOLTZMANN * T)
        swap = y_bot < y_top
        y_top, y_bot = np.where(swap, y_bot, y_top), np.where(swap, y_top, y_bot)
        acc = np.zeros((dim + 1, nwave), f8)
        for n in range(1, 200):
            acc += D * analyt_planck(n, y_bot, y_top)
    acc = np.where(T > 0.01, acc, 0.0)
    return (acc / deltalambda[None, :]).reshape(-1)


def corr_inc_energy(planck_grid, starflux, deltalambda, realstar, nwave, Tstar, dim):
    """K:420-468; returns (planck_grid, starflux, corr_factor) with the correction applied."""
    planck_grid = np.array(planck_grid, f8)
    starflux = None if starflux is None else np.array(starflux, f8)
    num = 0.0
    if realstar == 1:
        for xl in range(nwave):
            num += deltalambda[xl] * starflux[xl]
    else:
        for xl in range(nwave):
            num += deltalambda[xl] * PI * planck_grid[xl + dim * nwave]
    theo = STEFANBOLTZMANN * float(Tstar) ** 4.0
    corr = theo / num
    if realstar == 1:
        starflux *= corr
    else:
        planck_grid[dim * nwave:(dim + 1) * nwave] *= corr
    return planck_grid, starflux, corr


# ------------------------------------------------------------------------------- interpolation
def temp_inter(tlay, nint):
    """K:496-520"""
    tlay = np.asarray(tlay, f8)
    tint = np.empty(nint, f8)
    i = np.arange(1, nint - 1)
    tint[i] = tlay[i - 1] + 0.5 * (tlay[i] - tlay[i - 1])
    tint[0] = tlay[0] - 0.5 * (tlay[1] - tlay[0])
    tint[nint - 1] = tlay[nint - 2] + 0.5 * (tlay[nint - 2] - tlay[nint - 3])
    return tint


def _planck_rows(temp_vals, planck_grid, nwave, dim, step):
    g = np.asarray(planck_grid, f8).reshape(dim + 1, nwave)
    t = (np.asarray(temp_vals, f8) - 1.0) / step
    t = np.maximum(0.001, np.minimum(dim - 1.001, t))
    td = np.floor(t).astype(int)
    tu = np.ceil(t).astype(int)
    lin = g[td] * (tu - t)[:, None] + g[tu] * (t - td)[:, None]
    return np.where((td != tu)[:, None], lin, g[td])  # [row, x]


def planck_interpol_layer(temp, planck_grid, starflux, realstar, numlayers, nwave, dim, step):
    """K:923-977 -> planckband_lay flat [x][i], i in [0, numlayers+2)"""
    temp = np.asarray(temp, f8)
    out = np.zeros((nwave, numlayers + 2), f8)
    out[:, :numlayers] = _planck_rows(temp[:numlayers], planck_grid, nwave, dim, step).T
    if realstar == 1:
        out[:, numlayers] = np.asarray(starflux, f8) / PI
    else:
        out[:, numlayers] = np.asarray(planck_grid, f8)[dim * nwave:(dim + 1) * nwave]
    out[:, numlayers + 1] = _planck_rows(temp[numlayers:numlayers + 1], planck_grid, nwave, dim, step)[0]
    return out.reshape(-1)


def planck_interpol_interface(temp, planck_grid, numinterfaces, nwave, dim, step):
    """K:981-1011 -> planckband_int flat [x][i]"""
    rows = _planck_rows(np.asarray(temp, f8)[:numinterfaces], planck_grid, nwave, dim, step)
    return rows.T.copy().reshape(-1)


def pt_box(temp, press, gtemp, gpress, clamp="open", log_t=False):
    """index arithmetic shared by K:545-559, 665-679, 722-737, 777-791, 3228-3241"""
    temp, press, gtemp, gpress = (np.asarray(a, f8) for a in (temp, press, gtemp, gpress))
    nt, npr = gtemp.size, gpress.size
    if log_t:
        dT = (np.log10(gtemp[nt - 1]) - np.log10(gtemp[0])) / (nt - 1.0)
        t = (np.log10(temp) - np.log10(gtemp[0])) / dT
    else:
        dT = (gtemp[nt - 1] - gtemp[0]) / (nt - 1.0)
        t = (temp - gtemp[0]) / dT
    dP = (np.log10(gpress[npr - 1]) - np.log10(gpress[0])) / (npr - 1.0)
    p = (np.log10(press) - np.log10(gpress[0])) / dP
    if clamp == "open":
        t = np.minimum(nt - 1.001, np.maximum(0.001, t))
        p = np.minimum(npr - 1.001, np.maximum(0.001, p))
    else:
        t = np.minimum(nt - 1.0, np.maximum(0.0, t))
        p = np.minimum(npr - 1.0, np.maximum(0.0, p))
    return p, t, np.floor(p).astype(int), np.ceil(p).astype(int), np.floor(t).astype(int), np.ceil(t).astype(int)


def _bilin(dd, ud, du, uu, p, t, pd, pu, td, tu):
    """four-branch bilinear form, K:561-608 / K:613-645; leading axis = layer"""
    sh = (-1,) + (1,) * (dd.ndim - 1)
    p, t = p.reshape(sh), t.reshape(sh)
    pdn, pup, tdn, tup = (a.reshape(sh).astype(f8) for a in (pd, pu, td, tu))
    full = dd * (pup - p) * (tup - t) + ud * (p - pdn) * (tup - t) + du * (pup - p) * (t - tdn) + uu * (p - pdn) * (t - tdn)
    only_p = dd * (pup - p) + ud * (p - pdn)
    only_t = dd * (tup - t) + du * (t - tdn)
    pe, te = (pdn == pup), (tdn == tup)
    return np.where(~pe & ~te, full, np.where(te & ~pe, only_p, np.where(pe & ~te, only_t, dd)))


def _table_interp(table, rowlen, npress, ntemp, p, t, pd, pu, td, tu):
    tab = np.asarray(table, f8).reshape(ntemp, npress, rowlen)
    return _bilin(tab[td, pd], tab[td, pu], tab[tu, pd], tab[tu, pu], p, t, pd, pu, td, tu)


def opac_interpol(temp, opactemp, press, opacpress, ktable, crosstable, npress, ntemp, ny, nbin, n_i):
    """K:524-610 -> (opac flat [i][x][y], scat_cross flat [i][x])"""
    box = pt_box(np.asarray(temp)[:n_i], np.asarray(press)[:n_i], opactemp, opacpress, "open")
    opac = _table_interp(ktable, ny * nbin, npress, ntemp, *box)
    scat = _table_interp(crosstable, nbin, npress, ntemp, *box)
    return opac.reshape(-1), scat.reshape(-1)


def opac_species_interpol(temp, opactemp, press, opacpress, pretab, npress, ntemp, ny, nbin, n_i):
    """K:3209-3259"""
    box = pt_box(np.asarray(temp)[:n_i], np.asarray(press)[:n_i], opactemp, opacpress, "closed")
    return _table_interp(pretab, ny * nbin, npress, ntemp, *box).reshape(-1)


def scalar_interpol(temp, gtemp, press, gpress, tab, npress, ntemp, n, log_t=False):
    """K:649-699 (meanmolmass), 703-757 (kappa), 761-811 (cp, log T), 815-865 (entropy, log T),
    869-919 (phase number); table layout tab[p + npress*t]"""
    p, t, pd, pu, td, tu = pt_box(np.asarray(temp)[:n], np.asarray(press)[:n], gtemp, gpress, "open", log_t)
    tb = np.asarray(tab, f8).reshape(ntemp, npress)
    return _bilin(tb[td, pd], tb[td, pu], tb[tu, pd], tb[tu, pu], p, t, pd, pu, td, tu)


# ------------------------------------------------------------------------------- species mixing
def add_to_mixed_opac(vmr, opac_spec, opac_wg, meanmolmass, gauss_weight, gauss_y, mass_spec, s, ro_method,
                      ny, nbin, n_i):
    """K:3263-3399 with sort_array K:3152-3171.  Returns the updated opac_wg (flat [i][x][y])."""
    vmr = np.asarray(vmr, f8)[:n_i]
    mmm = np.asarray(meanmolmass, f8)[:n_i]
    mixed = np.array(opac_wg, f8)[:n_i * nbin * ny].reshape(n_i, nbin, ny)
    spec = np.asarray(opac_spec, f8)[:n_i * nbin * ny].reshape(n_i, nbin, ny)
    new = (vmr * mass_spec / mmm)[:, None, None] * spec
    negligible = (0.01 * mixed[..., 0] > new[..., ny - 1]) | (0.01 * new[..., 0] > mixed[..., ny - 1])
    corrk = negligible | (ro_method == 0) | (s == 0) | (ny == 1)
    out = np.where(corrk[..., None], mixed + new, mixed)
    idx = np.argwhere(~corrk)
    if idx.size == 0:
        return out.reshape(-1)
    assert ny == 20, "random overlap is hard-wired to 20 Gauss points in the reference (K:3314)"
    gw = np.asarray(gauss_weight, f8)
    gy = np.asarray(gauss_y, f8)
    M = mixed[idx[:, 0], idx[:, 1]]  # [c, 20]
    Nw = new[idx[:, 0], idx[:, 1]]
    nc = M.shape[0]
    # y_intersect: last y whose ordering differs from y-1 (K:3321-3329)
    gt = M > Nw
    flip = gt[:, 1:] != gt[:, :-1]
    yi = np.where(flip.any(axis=1), ny - 1 - np.argmax(flip[:, ::-1], axis=1), ny)
    outer = M[:, 0] > Nw[:, 0]
    # slot of pair (y1, y2) in the reference's scratch layout (K:3332-3365)
    y1 = np.arange(ny)[None, :, None]
    y2 = np.arange(ny)[None, None, :]
    yi3 = yi[:, None, None]
    posA = np.where(y2 < yi3, y2 + yi3 * y1, y1 + ny * y2)
    posB = np.where(y1 < yi3, y1 + yi3 * y2, y2 + ny * y1)
    pos = np.where(outer[:, None, None], posA, posB).reshape(nc, ny * ny)
    ksum = (M[:, :, None] + Nw[:, None, :]).reshape(nc, ny * ny)
    wsum = ((0.5 * gw)[:, None] * (0.5 * gw)[None, :]).reshape(1, ny * ny).repeat(nc, axis=0)
    K = np.empty_like(ksum)
    W = np.empty_like(wsum)
    rows = np.arange(nc)[:, None]
    K[rows, pos] = ksum
    W[rows, pos] = wsum
    # the exchange sort of K:3152-3171 swaps neighbours only on strict '<', i.e. it is a stable sort
    order = np.argsort(K, axis=1, kind="stable")
    K = np.take_along_axis(K, order, axis=1)
    W = np.take_along_axis(W, order, axis=1)
    n2 = ny * ny
    yg = np.empty_like(K)
    yg[:, 0] = 0.5 * W[:, 0]
    for w in range(1, n2):
        yg[:, w] = yg[:, w - 1] + 0.5 * W[:, w - 1] + 0.5 * W[:, w]
    # rebinning (K:3379-3396)
    res = M.copy()
    ycur = np.zeros(nc, int)
    done = np.zeros(nc, bool)
    r = np.arange(nc)
    for w in range(1, n2):
        g = gy[ycur]
        hit = (~done) & (yg[:, w] > g)
        if hit.any():
            val = (K[:, w - 1] * (yg[:, w] - g) + K[:, w] * (g - yg[:, w - 1])) / (yg[:, w] - yg[:, w - 1])
            res[r[hit], ycur[hit]] = val[hit]
            last = hit & (ycur >= ny - 1)
            done |= last
            adv = hit & ~last
            ycur[adv] += 1
        if done.all():
            break
    out[idx[:, 0], idx[:, 1]] = res
    return out.reshape(-1)


def calc_h2o_scat(temp, press, wave, vmr, mass_h2o, nbin, n_i):
    """K:3404-3440 with calc_index_h2o K:3174-3205 -> flat [i][x]"""
    T = np.asarray(temp, f8)[:n_i, None]
    P = np.asarray(press, f8)[:n_i, None]
    f = np.asarray(vmr, f8)[:n_i, None]
    wv = np.asarray(wave, f8)[None, :nbin]
    dens = f * P * mass_h2o / (KBOLTZMANN * T)
    lam = wv / 0.589e-4
    delta = np.minimum(1.0, dens) / 1.0
    theta = T / 273.15
    lUV, lIR = 0.229202, 5.432937
    a0, a1, a2, a3 = 0.244257733, 0.974634476e-2, -0.373234996e-2, 0.268678472e-3
    a4, a5, a6, a7 = 0.158920570e-2, 0.245934259e-2, 0.900704920, -0.166626219e-1
    A = delta * (a0 + a1 * delta + a2 * theta + a3 * lam ** 2 * theta + a4 * lam ** -2.0
                 + a5 / (lam ** 2 - lUV ** 2) + a6 / (lam ** 2 - lIR ** 2) + a7 * delta ** 2)
    index = ((2.0 * A + 1.0) / (1.0 - A)) ** 0.5
    n_ref = f * P / (KBOLTZMANN * T)
    King = (6.0 + 3.0 * 3e-4) / (6.0 - 7.0 * 3e-4)
    sc = 24.0 * PI ** 3.0 / (n_ref ** 2 * wv ** 4) * ((index ** 2 - 1.0) / (index ** 2 + 2.0)) ** 2 * King
    return np.where(wv < 2.5e-4, sc, 0.0).reshape(-1)


def add_to_mixed_scat(vmr, scat_spec, scat, nbin, n_i):
    """K:3444-3459"""
    out = np.array(scat, f8)[:n_i * nbin].reshape(n_i, nbin)
    out += np.asarray(vmr, f8)[:n_i, None] * np.asarray(scat_spec, f8)[:n_i * nbin].reshape(n_i, nbin)
    return out.reshape(-1)


def calc_total_g_0(scat, g0_clouds, scat_clouds, g_0):
    """K:472-492"""
    scat, g0_clouds, scat_clouds = (np.asarray(a, f8) for a in (scat, g0_clouds, scat_clouds))
    with np.errstate(invalid="ignore", divide="ignore"):
        return (g_0 * scat + g0_clouds * scat_clouds) / (scat + scat_clouds)


# ------------------------------------------------------------------------------- transmission
def E_parameter(w0, g0, i2s):
    """K:109-124"""
    E = np.maximum(1.0, 1.225 - 0.1582 * g0 - 0.1777 * w0 - 0.07465 * g0 ** 2.0 + 0.2351 * w0 * g0 - 0.05582 * w0 ** 2.0)
    return np.where((w0 > i2s) & (g0 >= 0), E, 1.0)


def _G_limiter(G):
    """K:218-231"""
    with np.errstate(invalid="ignore", divide="ignore"):
        return np.where(np.abs(G) < 1e8, G, 1e8 * G / np.abs(G))


def _cell_coeffs(ray, csc, cab, opac, mmm, dcol, dtc, g0, epsi, epsi2, mu_star, w_0_limit, scat_corr, i2s):
    """K:1076-1099 with K:128-290; all array arguments broadcast to [i][x][y]"""
    with np.errstate(invalid="ignore", divide="ignore", over="ignore"):
        w0 = np.minimum((ray + csc) / ((ray + csc) + (opac * mmm + cab)), w_0_limit)
        dtau = dcol * (opac + ray / mmm)
        del_tau = dtau + dtc
        E = E_parameter(w0, g0, i2s) if scat_corr == 1 else np.ones_like(w0)
        trans = np.exp(-1.0 / epsi * np.sqrt(E * (1.0 - w0 * g0) * (E - w0)) * del_tau)
        root = np.sqrt((E - w0) / (E * (1.0 - w0 * g0)))
        zm = 0.5 * (1.0 - root)
        zp = 0.5 * (1.0 + root)
        M = (zm * zm) * (trans * trans) - (zp * zp)
        N = zp * zm * (1.0 - (trans * trans))
        P = ((zm * zm) - (zp * zp)) * trans
        num = w0 * (E * (1.0 - w0 * g0) + g0 * epsi / epsi2)
        denom = E * epsi ** -2.0 * (E - w0) * (1.0 - w0 * g0) - mu_star ** -2.0
        second_p = 1.0 / epsi + 1.0 / (mu_star * E * (1.0 - w0 * g0))
        second_m = 1.0 / epsi - 1.0 / (mu_star * E * (1.0 - w0 * g0))
        third = epsi * w0 * g0 * mu_star / (epsi2 * E * (1.0 - w0 * g0))
        Gp = _G_limiter(0.5 * (num / denom * second_p + third))
        Gm = _G_limiter(0.5 * (num / denom * second_m - third))
    return dict(w0=w0, dtau=dtau, trans=trans, M=M, N=N, P=P, Gp=Gp, Gm=Gm)


def calc_trans_iso(delta_colmass, opac_wg_lay, meanmolmass_lay, scat_cross_lay, abs_cross_cl, scat_cross_cl,
                   g_0_tot_lay, g_0, epsi, epsi2, mu_star, w_0_limit, w_0_scat_limit, scat, nbin, ny, nlayer,
                   clouds, scat_corr, i2s):
    """K:1015-1104.  Returns dict of flat arrays: trans, dtau, M, N, P, Gp, Gm, w0 ([i][x][y]),
    dtau_clouds ([i][x]) and scat_trigger ([x][y], int32; 1 where set -- the caller zeroes it first)."""
    sh3 = (nlayer, nbin, ny)
    opac = np.asarray(opac_wg_lay, f8)[:nlayer * nbin * ny].reshape(sh3)
    mmm = np.asarray(meanmolmass_lay, f8)[:nlayer, None, None]
    dcol = np.asarray(delta_colmass, f8)[:nlayer, None, None]
    band = lambda a: np.asarray(a, f8)[:nlayer * nbin].reshape(nlayer, nbin, 1)
    g0 = band(g_0_tot_lay) if clouds == 1 else g_0
    ray = band(scat_cross_lay) if scat == 1 else 0.0
    csc = band(scat_cross_cl) if scat == 1 else 0.0
    cab = band(abs_cross_cl)
    dtc = dcol * (cab + csc) / mmm
    c = _cell_coeffs(ray, csc, cab, opac, mmm, dcol, dtc, g0, epsi, epsi2, mu_star, w_0_limit, scat_corr, i2s)
    out = {k: np.broadcast_to(v, sh3).reshape(-1).copy() for k, v in c.items()}
    out["dtau_clouds"] = np.broadcast_to(dtc, (nlayer, nbin, 1)).reshape(-1).copy()
    out["scat_trigger"] = (c["w0"] > w_0_scat_limit).any(axis=0).astype(np.int32).reshape(-1)
    return out


def calc_trans_noniso(delta_col_upper, delta_col_lower, opac_wg_lay, opac_wg_int, meanmolmass_lay,
                      meanmolmass_int, scat_cross_lay, scat_cross_int, abs_cl_lay, abs_cl_int, scat_cl_lay,
                      scat_cl_int, g_0_tot_lay, g_0_tot_int, g_0, epsi, epsi2, mu_star, w_0_limit,
                      w_0_scat_limit, scat, nbin, ny, nlayer, clouds, scat_corr, i2s):
    """K:1107-1243.  Returns (upper, lower, scat_trigger) with the dict layout of calc_trans_iso."""
    nint = nlayer + 1
    sh3 = (nlayer, nbin, ny)
    k_lay = np.asarray(opac_wg_lay, f8)[:nlayer * nbin * ny].reshape(sh3)
    k_int = np.asarray(opac_wg_int, f8)[:nint * nbin * ny].reshape(nint, nbin, ny)
    mm_lay = np.asarray(meanmolmass_lay, f8)[:nlayer]
    mm_int = np.asarray(meanmolmass_int, f8)[:nint]
    bl = lambda a: np.asarray(a, f8)[:nlayer * nbin].reshape(nlayer, nbin, 1)
    bi = lambda a: np.asarray(a, f8)[:nint * nbin].reshape(nint, nbin, 1)

    def halves(lay, inter):
        return (lay + inter[1:]) / 2.0, (inter[:-1] + lay) / 2.0

    if clouds == 1:
        g0_up, g0_low = halves(bl(g_0_tot_lay), bi(g_0_tot_int))
    else:
        g0_up = g0_low = g_0
    if scat == 1:
        ray_up, ray_low = halves(bl(scat_cross_lay), bi(scat_cross_int))
        csc_up, csc_low = halves(bl(scat_cl_lay), bi(scat_cl_int))
    else:
        ray_up = ray_low = csc_up = csc_low = 0.0
    cab_up, cab_low = halves(bl(abs_cl_lay), bi(abs_cl_int))
    opac_up, opac_low = halves(k_lay, k_int)
    mmm_up = ((mm_lay + mm_int[1:]) / 2.0)[:, None, None]
    mmm_low = ((mm_int[:-1] + mm_lay) / 2.0)[:, None, None]
    dcu = np.asarray(delta_col_upper, f8)[:nlayer, None, None]
    dcl = np.asarray(delta_col_lower, f8)[:nlayer, None, None]
    dtc_up = dcu * (cab_up + csc_up) / mmm_up
    dtc_low = dcl * (cab_low + csc_low) / mmm_low
    cu = _cell_coeffs(ray_up, csc_up, cab_up, opac_up, mmm_up, dcu, dtc_up, g0_up, epsi, epsi2, mu_star, w_0_limit, scat_corr, i2s)
    cl = _cell_coeffs(ray_low, csc_low, cab_low, opac_low, mmm_low, dcl, dtc_low, g0_low, epsi, epsi2, mu_star, w_0_limit, scat_corr, i2s)
    up = {k: np.broadcast_to(v, sh3).reshape(-1).copy() for k, v in cu.items()}
    low = {k: np.broadcast_to(v, sh3).reshape(-1).copy() for k, v in cl.items()}
    up["dtau_clouds"] = np.broadcast_to(dtc_up, (nlayer, nbin, 1)).reshape(-1).copy()
    low["dtau_clouds"] = np.broadcast_to(dtc_low, (nlayer, nbin, 1)).reshape(-1).copy()
    trig = ((cu["w0"] > w_0_scat_limit) | (cl["w0"] > w_0_scat_limit)).any(axis=0).astype(np.int32).reshape(-1)
    return up, low, trig


def calc_delta_z(tlay, pint, meanmolmass_lay, g, nlayer):
    """K:1247-1261"""
    tlay, pint, mmm = (np.asarray(a, f8) for a in (tlay, pint, meanmolmass_lay))
    return KBOLTZMANN * tlay[:nlayer] / (mmm[:nlayer] * g) * np.log(pint[:nlayer] / pint[1:nlayer + 1])


def fdir(planckband_lay, dtau_a, dtau_b, z_lay, mu_star, R_planet, R_star, a, dir_beam, geom, nint, nbin, ny):
    """K:1265-1309 (dtau_b None) / K:1313-1362 -> (F_dir, Fc_dir or None), flat [i][x][y].
    Fc_dir at the top interface is never written by the reference and is returned as 0."""
    nlay = nint - 1
    noniso = dtau_b is not None
    B = np.asarray(planckband_lay, f8).reshape(nbin, nlay + 2)
    I_dir = ((R_star / a) * (R_star / a)) * PI * B[:, nlay]
    F_toa = np.repeat(-dir_beam * mu_star * I_dir, ny)  # [col]
    da = np.asarray(dtau_a, f8)[:nlay * nbin * ny].reshape(nlay, nbin * ny)
    db = np.asarray(dtau_b, f8)[:nlay * nbin * ny].reshape(nlay, nbin * ny) if noniso else None
    F = np.empty((nint, nbin * ny), f8)
    Fc = np.zeros((nint, nbin * ny), f8) if noniso else None
    F[nlay] = F_toa
    with np.errstate(over="ignore", invalid="ignore", under="ignore"):
        for i in range(nlay - 1, -1, -1):
            cur = F_toa.copy()
            curc = None
            for j in range(nlay - 1, i - 1, -1):
                if geom == 1:
                    mu_j = -np.sqrt(1.0 - ((R_planet + z_lay[i]) / (R_planet + z_lay[j])) ** 2.0 * (1.0 - mu_star ** 2.0))
                else:
                    mu_j = mu_star
                if noniso:
                    curc = cur * np.exp(da[j] / mu_j)
                    cur = cur * np.exp((da[j] + db[j]) / mu_j)
                else:
                    cur = cur * np.exp(da[j] / mu_j)
            F[i] = cur
            if noniso:
                Fc[i] = curc
    return F.reshape(-1), (Fc.reshape(-1) if noniso else None)


# ------------------------------------------------------------------------------- flux sweeps
def _tiny_abs(F):
    return np.where(np.abs(F) < 1e-100, np.abs(F), F)


def fband_iso(F_down, F_up, F_dir, planckband_lay, w_0, M_term, N_term, P_term, G_plus, G_minus, surf_albedo,
              g_0_tot_lay, g_0, Rstar, a, nint, nbin, f_factor, mu_star, ny, epsi, dir_beam, clouds, scat_corr,
              i2s, npass=1):
    """K:1366-1517, repeated npass times (computation.py:537).  Returns new (F_down, F_up)."""
    nlay = nint - 1
    nc = nbin * ny
    Fd = np.array(F_down, f8).reshape(nint, nc)
    Fu = np.array(F_up, f8).reshape(nint, nc)
    Fdir = np.asarray(F_dir, f8).reshape(nint, nc)
    B = np.repeat(np.asarray(planckband_lay, f8).reshape(nbin, nlay + 2), ny, axis=0)  # [col, i]
    c3 = lambda v: np.asarray(v, f8)[:nlay * nc].reshape(nlay, nc)
    w0a, Ma, Na, Pa, Gpa, Gma = (c3(v) for v in (w_0, M_term, N_term, P_term, G_plus, G_minus))
    alb = np.repeat(np.asarray(surf_albedo, f8), ny)
    g0a = np.repeat(np.asarray(g_0_tot_lay, f8)[:nlay * nbin].reshape(nlay, nbin), ny, axis=1) if clouds == 1 else None
    with np.errstate(invalid="ignore", divide="ignore", over="ignore"):
        for _ in range(npass):
            w0 = E = None
            for i in range(nint - 1, -1, -1):
                if i == nint - 1:
                    Fd[i] = (1.0 - dir_beam) * f_factor * ((Rstar / a) * (Rstar / a)) * PI * B[:, i]
                    continue
                w0, M, N, P, Gp, Gm = w0a[i], Ma[i], Na[i], Pa[i], Gpa[i], Gma[i]
                g0 = g0a[i] if clouds == 1 else g_0
                E = E_parameter(w0, g0, i2s) if scat_corr == 1 else np.ones(nc)
                flux_terms = P * Fd[i + 1] - N * Fu[i]
                planck_terms = B[:, i] * (N + M - P)
                direct = Fdir[i] / (-mu_star) * (Gm * M + Gp * N) - Fdir[i + 1] / (-mu_star) * P * Gm
                direct = np.fmin(0.0, direct)
                Fd[i] = _tiny_abs(1.0 / M * (flux_terms + 2.0 * PI * epsi * (1.0 - w0) / (E - w0) * planck_terms + direct))
            for i in range(nint):
                if i == 0:
                    refl = alb * (Fdir[0] + Fd[0])
                    boa = (1.0 - alb) * PI * (1.0 - w0) / (E - w0) * B[:, nint]
                    Fu[0] = refl + boa
                    continue
                w0, M, N, P, Gp, Gm = w0a[i - 1], Ma[i - 1], Na[i - 1], Pa[i - 1], Gpa[i - 1], Gma[i - 1]
                g0 = g0a[i - 1] if clouds == 1 else g_0
                E = E_parameter(w0, g0, i2s) if scat_corr == 1 else np.ones(nc)
                flux_terms = P * Fu[i - 1] - N * Fd[i]
                planck_terms = B[:, i - 1] * (N + M - P)
                direct = Fdir[i] / (-mu_star) * (Gm * N + Gp * M) - Fdir[i - 1] / (-mu_star) * P * Gp
                direct = np.fmin(0.0, direct)
                Fu[i] = _tiny_abs(1.0 / M * (flux_terms + 2.0 * PI * epsi * (1.0 - w0) / (E - w0) * planck_terms + direct))
    return Fd.reshape(-1), Fu.reshape(-1)


def fband_noniso(F_down, F_up, Fc_down, Fc_up, F_dir, Fc_dir, planckband_lay, planckband_int, up, low,
                 surf_albedo, g_0_tot_lay, g_0_tot_int, g_0, Rstar, a, nint, nbin, f_factor, mu_star, ny, epsi,
                 delta_tau_limit, dir_beam, clouds, scat_corr, i2s, npass=1):
    """K:1521-1799.  `up` / `low` are dicts with flat arrays w0, dtau, dtau_clouds, M, N, P, Gp, Gm.
    Returns new (F_down, F_up, Fc_down, Fc_up)."""
    nlay = nint - 1
    nc = nbin * ny
    Fd = np.array(F_down, f8).reshape(nint, nc)
    Fu = np.array(F_up, f8).reshape(nint, nc)
    Fcd = np.array(Fc_down, f8).reshape(nint, nc)
    Fcu = np.array(Fc_up, f8).reshape(nint, nc)
    Fdir = np.asarray(F_dir, f8).reshape(nint, nc)
    Fcdir = np.asarray(Fc_dir, f8).reshape(nint, nc)
    BL = np.repeat(np.asarray(planckband_lay, f8).reshape(nbin, nlay + 2), ny, axis=0)
    BI = np.repeat(np.asarray(planckband_int, f8).reshape(nbin, nint), ny, axis=0)
    c3 = lambda v: np.asarray(v, f8)[:nlay * nc].reshape(nlay, nc)
    cb = lambda v: np.repeat(np.asarray(v, f8)[:nlay * nbin].reshape(nlay, nbin), ny, axis=1)
    U = {k: c3(up[k]) for k in ("w0", "dtau", "M", "N", "P", "Gp", "Gm")}
    L = {k: c3(low[k]) for k in ("w0", "dtau", "M", "N", "P", "Gp", "Gm")}
    U["dt"] = U["dtau"] + cb(up["dtau_clouds"])
    L["dt"] = L["dtau"] + cb(low["dtau_clouds"])
    alb = np.repeat(np.asarray(surf_albedo, f8), ny)
    if clouds == 1:
        gl = np.repeat(np.asarray(g_0_tot_lay, f8)[:nlay * nbin].reshape(nlay, nbin), ny, axis=1)
        gi = np.repeat(np.asarray(g_0_tot_int, f8)[:nint * nbin].reshape(nint, nbin), ny, axis=1)
        g0U = (gl + gi[1:]) / 2.0
        g0L = (gi[:-1] + gl) / 2.0
    tpe = 2.0 * PI * epsi

    def half(H, g0H, k):
        w0 = H["w0"][k]
        g0 = g0H[k] if clouds == 1 else g_0
        E = E_parameter(w0, g0, i2s) if scat_corr == 1 else np.ones(nc)
        return w0, H["dt"][k], H["M"][k], H["N"][k], H["P"][k], H["Gp"][k], H["Gm"][k], g0, E

    with np.errstate(invalid="ignore", divide="ignore", over="ignore"):
        for _ in range(npass):
            w0l = El = None
            for i in range(nint - 1, -1, -1):
                if i == nint - 1:
                    Fd[i] = (1.0 - dir_beam) * f_factor * ((Rstar / a) * (Rstar / a)) * PI * BL[:, i]
                    continue
                w0u, dtu, Mu, Nu, Pu, Gpu, Gmu, g0u, Eu = half(U, g0U if clouds == 1 else None, i)
                w0l, dtl, Ml, Nl, Pl, Gpl, Gml, g0l, El = half(L, g0L if clouds == 1 else None, i)
                # upper half (K:1640-1664)
                iso_pt = (BI[:, i + 1] + BL[:, i]) / 2.0 * (Nu + Mu - Pu)
                pgrad = (BL[:, i] - BI[:, i + 1]) / dtu
                non_pt = BL[:, i] * (Mu + Nu) - BI[:, i + 1] * Pu + epsi / (Eu * (1.0 - w0u * g0u)) * (Pu - Mu + Nu) * pgrad
                pt = np.where(dtu < delta_tau_limit, iso_pt, non_pt)
                ft = Pu * Fd[i + 1] - Nu * Fcu[i]
                dr = Fcdir[i] / (-mu_star) * (Gmu * Mu + Gpu * Nu) - Fdir[i + 1] / (-mu_star) * Gmu * Pu
                dr = np.fmin(0.0, dr)
                Fcd[i] = _tiny_abs(1.0 / Mu * (ft + tpe * (1.0 - w0u) / (Eu - w0u) * pt + dr))
                # lower half (K:1667-1691)
                iso_pt = (BI[:, i] + BL[:, i]) / 2.0 * (Nl + Ml - Pl)
                pgrad = (BI[:, i] - BL[:, i]) / dtl
                non_pt = BI[:, i] * (Ml + Nl) - BL[:, i] * Pl + epsi / (El * (1.0 - w0l * g0l)) * (Pl - Ml + Nl) * pgrad
                pt = np.where(dtl < delta_tau_limit, iso_pt, non_pt)
                ft = Pl * Fcd[i] - Nl * Fu[i]
                dr = Fdir[i] / (-mu_star) * (Gml * Ml + Gpl * Nl) - Fcdir[i] / (-mu_star) * Pl * Gml
                dr = np.fmin(0.0, dr)
                Fd[i] = _tiny_abs(1.0 / Ml * (ft + tpe * (1.0 - w0l) / (El - w0l) * pt + dr))
            for i in range(nint):
                if i == 0:
                    refl = alb * (Fdir[0] + Fd[0])
                    boa = (1.0 - alb) * PI * (1.0 - w0l) / (El - w0l) * BL[:, nint]
                    Fu[0] = refl + boa
                    continue
                k = i - 1
                w0l, dtl, Ml, Nl, Pl, Gpl, Gml, g0l, El = half(L, g0L if clouds == 1 else None, k)
                w0u, dtu, Mu, Nu, Pu, Gpu, Gmu, g0u, Eu = half(U, g0U if clouds == 1 else None, k)
                # lower half (K:1744-1768)
                iso_pt = (BI[:, k] + BL[:, k]) / 2.0 * (Nl + Ml - Pl)
                pgrad = (BI[:, k] - BL[:, k]) / dtl
                non_pt = BL[:, k] * (Ml + Nl) - BI[:, k] * Pl + epsi / (El * (1.0 - w0l * g0l)) * pgrad * (Ml - Pl - Nl)
                pt = np.where(dtl < delta_tau_limit, iso_pt, non_pt)
                ft = Pl * Fu[k] - Nl * Fcd[k]
                dr = Fcdir[k] / (-mu_star) * (Gml * Nl + Gpl * Ml) - Fdir[k] / (-mu_star) * Pl * Gpl
                dr = np.fmin(0.0, dr)
                Fcu[k] = 1.0 / Ml * (ft + tpe * (1.0 - w0l) / (El - w0l) * pt + dr)
                # the reference's tiny-value clean-up addresses Fc_up[i], not [i-1] (K:1763)
                Fcu[i] = _tiny_abs(Fcu[i])
                # upper half (K:1771-1795)
                iso_pt = (BI[:, i] + BL[:, k]) / 2.0 * (Nu + Mu - Pu)
                pgrad = (BL[:, k] - BI[:, i]) / dtu
                non_pt = BI[:, i] * (Mu + Nu) - BL[:, k] * Pu + epsi / (Eu * (1.0 - w0u * g0u)) * pgrad * (Mu - Pu - Nu)
                pt = np.where(dtu < delta_tau_limit, iso_pt, non_pt)
                ft = Pu * Fcu[k] - Nu * Fd[i]
                dr = Fdir[i] / (-mu_star) * (Gmu * Nu + Gpu * Mu) - Fcdir[k] / (-mu_star) * Pu * Gpu
                dr = np.fmin(0.0, dr)
                Fu[i] = _tiny_abs(1.0 / Mu * (ft + tpe * (1.0 - w0u) / (Eu - w0u) * pt + dr))
    return Fd.reshape(-1), Fu.reshape(-1), Fcd.reshape(-1), Fcu.reshape(-1)


def _thomas(b0, d0, rows_b, rows_c, rows_d, d_top):
    """Thomas sweep exactly as K:1916-1967 / K:2217-2284: row 0 = (b0, c=1, d0); the reference uses the
    previous row's c as the sub-diagonal.  rows_*: lists over matrix rows 1..n-2 of [col] arrays."""
    c_i = np.ones_like(d0)
    cp = [c_i / b0]
    dp = [d0 / b0]
    for b, c, d in zip(rows_b, rows_c, rows_d):
        c_prev = c_i
        c_i = c
        den = b - c_prev * cp[-1]
        cp.append(c_i / den)
        dp.append((d - c_prev * dp[-1]) / den)
    x_last = (d_top - c_i * dp[-1]) / (0.0 - c_i * cp[-1])
    return cp, dp, x_last


def fband_matrix_iso(F_down, F_up, F_dir, planckband_lay, w_0, M_term, N_term, P_term, G_plus, G_minus,
                     g_0_tot_lay, scat_trigger, trans_wg, surf_albedo, g_0, Rstar, a, nint, nbin, f_factor,
                     mu_star, ny, epsi, dir_beam, clouds, scat_corr, i2s):
    """K:1803-2024.  Returns new (F_down, F_up)."""
    nlay = nint - 1
    nc = nbin * ny
    Fd = np.array(F_down, f8).reshape(nint, nc)
    Fu = np.array(F_up, f8).reshape(nint, nc)
    Fdir = np.asarray(F_dir, f8).reshape(nint, nc)
    B = np.repeat(np.asarray(planckband_lay, f8).reshape(nbin, nlay + 2), ny, axis=0)
    c3 = lambda v: np.asarray(v, f8)[:nlay * nc].reshape(nlay, nc)
    w0a, Ma, Na, Pa, Gpa, Gma, Ta = (c3(v) for v in (w_0, M_term, N_term, P_term, G_plus, G_minus, trans_wg))
    alb = np.repeat(np.asarray(surf_albedo, f8), ny)
    g0a = np.repeat(np.asarray(g_0_tot_lay, f8)[:nlay * nbin].reshape(nlay, nbin), ny, axis=1) if clouds == 1 else None
    trig = np.asarray(scat_trigger).reshape(nc) == 1
    toa = (1.0 - dir_beam) * f_factor * ((Rstar / a) * (Rstar / a)) * PI * B[:, nlay]
    with np.errstate(invalid="ignore", divide="ignore", over="ignore"):
        # --- matrix branch
        rb, rc, rd = [], [], []
        for j in range(nlay):
            w0, M, N, P, Gp, Gm = w0a[j], Ma[j], Na[j], Pa[j], Gpa[j], Gma[j]
            g0 = g0a[j] if clouds == 1 else g_0
            E = E_parameter(w0, g0, i2s) if scat_corr == 1 else np.ones(nc)
            alpha = P / M
            beta = -N / M
            pt = 2.0 * PI * epsi * (1.0 - w0) / (E - w0) * (N + M - P) * B[:, j]
            dd = np.fmin(0.0, Fdir[j] / (-mu_star) * (Gm * M + Gp * N) - Fdir[j + 1] / (-mu_star) * P * Gm)
            du = np.fmin(0.0, Fdir[j + 1] / (-mu_star) * (Gm * N + Gp * M) - Fdir[j] / (-mu_star) * P * Gp)
            sd = 1.0 / M * (pt + dd)
            su = 1.0 / M * (pt + du)
            rb += [-beta, -beta]
            rc += [-alpha, np.ones(nc)]
            rd += [sd, su]
        w0 = w0a[0]
        g0 = g0a[0] if clouds == 1 else g_0
        E = E_parameter(w0, g0, i2s) if scat_corr == 1 else np.ones(nc)
        src_boa = alb * Fdir[0] + (1.0 - alb) * PI * (1.0 - w0) / (E - w0) * B[:, nint]
        cp, dp, x = _thomas(-alb, src_boa, rb, rc, rd, toa)
        n_matrix = 2 * nint
        mFd = np.empty_like(Fd)
        mFu = np.empty_like(Fu)
        mFu[nint - 1] = x
        for i in range(n_matrix - 2, -1, -1):
            x = dp[i] - cp[i] * x
            if i % 2 == 0:
                mFd[i // 2] = x
            else:
                mFu[(i - 1) // 2] = x
        # --- pure absorption branch
        aFd = np.empty_like(Fd)
        aFu = np.empty_like(Fu)
        aFd[nlay] = toa
        for i in range(nlay - 1, -1, -1):
            aFd[i] = _tiny_abs(Ta[i] * aFd[i + 1] + 2.0 * PI * epsi * (1.0 - Ta[i]) * B[:, i])
        aFu[0] = alb * (Fdir[0] + aFd[0]) + (1.0 - alb) * PI * B[:, nint]
        for i in range(1, nint):
            aFu[i] = _tiny_abs(Ta[i - 1] * aFu[i - 1] + 2.0 * PI * epsi * (1.0 - Ta[i - 1]) * B[:, i - 1])
    Fd = np.where(trig[None, :], mFd, aFd)
    Fu = np.where(trig[None, :], mFu, aFu)
    return Fd.reshape(-1), Fu.reshape(-1)


def fband_matrix_noniso(F_down, F_up, Fc_down, Fc_up, F_dir, Fc_dir, planckband_lay, planckband_int, up, low,
                        g_0_tot_lay, g_0_tot_int, scat_trigger, surf_albedo, g_0, Rstar, a, nint, nbin, f_factor,
                        mu_star, ny, epsi, delta_tau_limit, dir_beam, clouds, scat_corr, i2s):
    """K:2028-2424.  `up`/`low` additionally carry `trans`.  Returns (F_down, F_up, Fc_down, Fc_up)."""
    nlay = nint - 1
    nc = nbin * ny
    Fcd0 = np.array(Fc_down, f8).reshape(nint, nc)
    Fcu0 = np.array(Fc_up, f8).reshape(nint, nc)
    Fdir = np.asarray(F_dir, f8).reshape(nint, nc)
    Fcdir = np.asarray(Fc_dir, f8).reshape(nint, nc)
    BL = np.repeat(np.asarray(planckband_lay, f8).reshape(nbin, nlay + 2), ny, axis=0)
    BI = np.repeat(np.asarray(planckband_int, f8).reshape(nbin, nint), ny, axis=0)
    c3 = lambda v: np.asarray(v, f8)[:nlay * nc].reshape(nlay, nc)
    cb = lambda v: np.repeat(np.asarray(v, f8)[:nlay * nbin].reshape(nlay, nbin), ny, axis=1)
    U = {k: c3(up[k]) for k in ("w0", "dtau", "M", "N", "P", "Gp", "Gm", "trans")}
    L = {k: c3(low[k]) for k in ("w0", "dtau", "M", "N", "P", "Gp", "Gm", "trans")}
    U["dt"] = U["dtau"] + cb(up["dtau_clouds"])
    L["dt"] = L["dtau"] + cb(low["dtau_clouds"])
    alb = np.repeat(np.asarray(surf_albedo, f8), ny)
    if clouds == 1:
        gl = np.repeat(np.asarray(g_0_tot_lay, f8)[:nlay * nbin].reshape(nlay, nbin), ny, axis=1)
        gi = np.repeat(np.asarray(g_0_tot_int, f8)[:nint * nbin].reshape(nint, nbin), ny, axis=1)
    trig = np.asarray(scat_trigger).reshape(nc) == 1
    toa = (1.0 - dir_beam) * f_factor * ((Rstar / a) * (Rstar / a)) * PI * BL[:, nlay]
    tpe = 2.0 * PI * epsi
    with np.errstate(invalid="ignore", divide="ignore", over="ignore"):
        rb, rc, rd = [], [], []
        for j in range(2 * nlay):
            k = j // 2
            if j % 2 == 0:  # lower half (K:2111-2149)
                H = L
                g0 = (gi[k] + gl[k]) / 2.0 if clouds == 1 else g_0
                Ba, Bb = BI[:, k], BL[:, k]
                w0, M, N, P, Gp, Gm, dt = H["w0"][k], H["M"][k], H["N"][k], H["P"][k], H["Gp"][k], H["Gm"][k], H["dt"][k]
                E = E_parameter(w0, g0, i2s) if scat_corr == 1 else np.ones(nc)
                iso = (N + M - P) * (Ba + Bb) / 2.0
                pgrad = (Ba - Bb) / dt
                ptd = (M + N) * Ba - P * Bb + epsi / (E * (1.0 - w0 * g0)) * (P - M + N) * pgrad
                ptu = (M + N) * Bb - P * Ba + epsi / (E * (1.0 - w0 * g0)) * (M - N - P) * pgrad
                dd = Fdir[k] / (-mu_star) * (Gm * M + Gp * N) - Fcdir[k] / (-mu_star) * P * Gm
                du = Fcdir[k] / (-mu_star) * (Gm * N + Gp * M) - Fdir[k] / (-mu_star) * P * Gp
            else:  # upper half (K:2150-2188)
                H = U
                g0 = (gi[k + 1] + gl[k]) / 2.0 if clouds == 1 else g_0
                Bl_, Bi_ = BL[:, k], BI[:, k + 1]
                w0, M, N, P, Gp, Gm, dt = H["w0"][k], H["M"][k], H["N"][k], H["P"][k], H["Gp"][k], H["Gm"][k], H["dt"][k]
                E = E_parameter(w0, g0, i2s) if scat_corr == 1 else np.ones(nc)
                iso = (N + M - P) * (Bl_ + Bi_) / 2.0
                pgrad = (Bl_ - Bi_) / dt
                ptd = (M + N) * Bl_ - P * Bi_ + epsi / (E * (1.0 - w0 * g0)) * (P - M + N) * pgrad
                ptu = (M + N) * Bi_ - P * Bl_ + epsi / (E * (1.0 - w0 * g0)) * (M - N - P) * pgrad
                dd = Fcdir[k] / (-mu_star) * (Gm * M + Gp * N) - Fdir[k + 1] / (-mu_star) * P * Gm
                du = Fdir[k + 1] / (-mu_star) * (Gm * N + Gp * M) - Fcdir[k] / (-mu_star) * P * Gp
            thin = dt < delta_tau_limit
            ptd = np.where(thin, iso, ptd)
            ptu = np.where(thin, iso, ptu)
            dd = np.fmin(0.0, dd)
            du = np.fmin(0.0, du)
            alpha = P / M
            beta = -N / M
            sd = 1.0 / M * (tpe * (1.0 - w0) / (E - w0) * ptd + dd)
            su = 1.0 / M * (tpe * (1.0 - w0) / (E - w0) * ptu + du)
            rb += [-beta, -beta]
            rc += [-alpha, np.ones(nc)]
            rd += [sd, su]
        # the last half-layer contributes its "down" row and its "up" row; n_matrix = 4*nint-2 = 4*nlay+2
        w0 = L["w0"][0]
        g0 = (gi[0] + gl[0]) / 2.0 if clouds == 1 else g_0
        E = E_parameter(w0, g0, i2s) if scat_corr == 1 else np.ones(nc)
        src_boa = alb * Fdir[0] + (1.0 - alb) * PI * (1.0 - w0) / (E - w0) * BL[:, nint]
        cp, dp, x = _thomas(-alb, src_boa, rb, rc, rd, toa)
        n_matrix = 4 * nint - 2
        mFd = np.zeros((nint, nc))
        mFu = np.zeros((nint, nc))
        mFcd = Fcd0.copy()
        mFcu = Fcu0.copy()
        mFu[nint - 1] = x
        for i in range(n_matrix - 2, -1, -1):
            x = dp[i] - cp[i] * x
            x = np.where(x < 1e-100, np.abs(x), x)
            r = i % 4
            if r == 0:
                mFd[i // 4] = x
            elif r == 1:
                mFu[(i - 1) // 4] = x
            elif r == 2:
                mFcd[(i - 2) // 4] = x
            else:
                mFcu[(i - 3) // 4] = x
        # --- pure absorption branch (K:2286-2422)
        aFd = np.empty((nint, nc))
        aFu = np.empty((nint, nc))
        aFcd = Fcd0.copy()
        aFcu = Fcu0.copy()
        aFd[nlay] = toa
        for i in range(nlay - 1, -1, -1):
            tu, tl, dtu, dtl = U["trans"][i], L["trans"][i], U["dt"][i], L["dt"][i]
            iso = (BI[:, i + 1] + BL[:, i]) / 2.0 * (1.0 - tu)
            pg = (BL[:, i] - BI[:, i + 1]) / dtu
            non = BL[:, i] - tu * BI[:, i + 1] + epsi * (tu - 1.0) * pg
            aFcd[i] = _tiny_abs(tu * aFd[i + 1] + tpe * np.where(dtu < delta_tau_limit, iso, non))
            iso = (BI[:, i] + BL[:, i]) / 2.0 * (1.0 - tl)
            pg = (BI[:, i] - BL[:, i]) / dtl
            non = BI[:, i] - tl * BL[:, i] + epsi * (tl - 1.0) * pg
            aFd[i] = _tiny_abs(tl * aFcd[i] + tpe * np.where(dtl < delta_tau_limit, iso, non))
        aFu[0] = alb * (Fdir[0] + aFd[0]) + (1.0 - alb) * PI * BL[:, nint]
        for i in range(1, nint):
            k = i - 1
            tu, tl, dtu, dtl = U["trans"][k], L["trans"][k], U["dt"][k], L["dt"][k]
            iso = (BI[:, k] + BL[:, k]) / 2.0 * (1.0 - tl)
            pg = (BI[:, k] - BL[:, k]) / dtl
            non = BL[:, k] - tl * BI[:, k] + epsi * pg * (1.0 - tl)
            aFcu[k] = tl * aFu[k] + tpe * np.where(dtl < delta_tau_limit, iso, non)
            aFcu[i] = _tiny_abs(aFcu[i])
            iso = (BI[:, i] + BL[:, k]) / 2.0 * (1.0 - tu)
            pg = (BL[:, k] - BI[:, i]) / dtu
            non = BI[:, i] - tu * BL[:, k] + epsi * pg * (1.0 - tu)
            aFu[i] = _tiny_abs(tu * aFcu[k] + tpe * np.where(dtu < delta_tau_limit, iso, non))
    sel = trig[None, :]
    return (np.where(sel, mFd, aFd).reshape(-1), np.where(sel, mFu, aFu).reshape(-1),
            np.where(sel, mFcd, aFcd).reshape(-1), np.where(sel, mFcu, aFcu).reshape(-1))


# ------------------------------------------------------------------------------- integration
def integrate_flux(deltalambda, F_down_wg, F_up_wg, F_dir_wg, gauss_weight, nbin, nint, ny):
    """K:2428-2513 with a fixed summation order (y ascending, then x ascending).
    Returns dict: F_down_band, F_up_band, F_dir_band ([i][x]), F_down_tot, F_up_tot, F_net ([i])."""
    gw = 0.5 * np.asarray(gauss_weight, f8)
    dl = np.asarray(deltalambda, f8)

    def band(Fwg):
        F = np.asarray(Fwg, f8)[:nint * nbin * ny].reshape(nint, nbin, ny)
        acc = np.zeros((nint, nbin), f8)
        for y in range(ny):
            acc += gw[y] * F[:, :, y]
        return acc

    dn, up, dr = band(F_down_wg), band(F_up_wg), band(F_dir_wg)
    up_tot = np.zeros(nint, f8)
    dn_tot = np.zeros(nint, f8)
    for x in range(nbin):
        up_tot += up[:, x] * dl[x]
        dn_tot += (dr[:, x] + dn[:, x]) * dl[x]
    return dict(F_down_band=dn.reshape(-1), F_up_band=up.reshape(-1), F_dir_band=dr.reshape(-1),
                F_down_tot=dn_tot, F_up_tot=up_tot, F_net=up_tot - dn_tot)


# ------------------------------------------------------------------------------- temperature steps
def rad_temp_iter(F_down_tot, F_net, tlay, play, pint, T_store, prefactor, F_add_heat_lay, F_add_heat_sum,
                  F_smooth, F_smooth_sum, c_p_lay, meanmolmass_lay, itervalue, foreplay, g, numlayers,
                  physical_tstep, local_limit, adapt_interval, smooth, dim, step, F_intern, no_atmo):
    """K:2606-2764.  Returns dict with the updated tlay, abrt, T_store, prefactor, F_net_diff, F_smooth,
    F_smooth_sum.  Neighbour temperatures are read before any update (race-free reading of K:2659)."""
    nl = numlayers
    T = np.array(tlay, f8)
    T_store = np.array(T_store, f8)
    pref = np.array(prefactor, f8)
    F_net = np.asarray(F_net, f8)
    Fsm = np.array(F_smooth, f8)
    Fsms = np.array(F_smooth_sum, f8)
    Fnd = F_net[:nl] - F_net[1:nl + 1] + np.asarray(F_add_heat_lay, f8)[:nl]
    if smooth == 1:
        t_mid = T[:nl].copy()
        for i in range(1, nl - 1):
            if play[i] < 1e6:
                t_mid[i] = (T[i - 1] + T[i + 1]) / 2.0
        Fsm[:nl] = (t_mid - T[:nl]) ** 7.0
        acc = 0.0
        for j in range(nl):
            acc += Fsm[j]
            Fsms[j] = acc
    comb = np.empty(nl + 1, f8)
    comb[:nl] = Fnd + Fsm[:nl]
    comb[nl] = F_intern - F_net[0]
    if abs(F_intern - F_net[1]) / (F_down_tot[nl] + F_intern) > 0.5 * local_limit:
        comb[nl] = F_intern - F_net[1]
    abrt = np.zeros(nl + 1, np.int32)
    T_old = T.copy()
    for i in range(nl + 1):
        c = comb[i]
        if physical_tstep == 0:
            if itervalue == foreplay:
                pref[i] = 1e0
            if itervalue == 10000:
                pref[i] = 1e-1
            delta_t = pref[i] * play[0] / abs(c) ** 0.9 if c != 0 else 0.0
            dT = c / (pint[0] - pint[1]) * delta_t
            if abs(dT) > 500.0:
                dT = 500.0 * c / abs(c)
            if itervalue % adapt_interval == 0:
                T_store[i] = T_old[i]
            if itervalue % adapt_interval == adapt_interval - 1:
                if abs(T_old[i] - T_store[i]) < adapt_interval / 2.0 * abs(dT):
                    pref[i] /= 1.5
                else:
                    pref[i] *= 1.1
        else:
            k = i if i < nl else 0
            dT = g / (c_p_lay[k] / (meanmolmass_lay[k] / AMU)) * c / (pint[k] - pint[k + 1]) * physical_tstep
        Tn = T_old[i] + dT
        if no_atmo == 1 and i != nl:
            Tn = 1.001
        T[i] = min(max(Tn, 1.001), dim * step - 1.001)
        if i < nl:
            ok = abs(F_intern + F_add_heat_sum[i] + Fsms[i] - F_net[i + 1]) / (F_down_tot[nl] + F_intern) < local_limit
        else:
            ok = abs(F_intern - F_net[0]) / (F_down_tot[nl] + F_intern) < local_limit
        abrt[i] = 1 if ok else 0
    return dict(tlay=T, abrt=abrt, T_store=T_store, prefactor=pref, F_net_diff=Fnd, F_smooth=Fsm, F_smooth_sum=Fsms)


def conv_temp_iter(F_net, tlay, play, pint, T_store, prefactor, marked_red, F_add_heat_lay, F_smooth,
                   F_smooth_sum, numlayers, itervalue, adapt_interval, smooth, F_intern):
    """K:2768-2884"""
    nl = numlayers
    T = np.array(tlay, f8)
    T_store = np.array(T_store, f8)
    pref = np.array(prefactor, f8)
    F_net = np.asarray(F_net, f8)
    Fsm = np.array(F_smooth, f8)
    Fsms = np.array(F_smooth_sum, f8)
    Fnd = F_net[:nl] - F_net[1:nl + 1] + np.asarray(F_add_heat_lay, f8)[:nl]
    if smooth == 1:
        t_mid = T[:nl].copy()
        for i in range(1, nl - 1):  # the reference lacks the i > 0 guard here (K:2808) and reads tlay[-1]
            if play[i] < 1e6:
                t_mid[i] = (T[i - 1] + T[i + 1]) / 2.0
        Fsm[:nl] = (t_mid - T[:nl]) ** 7.0
        acc = 0.0
        for j in range(nl):
            acc += Fsm[j]
            Fsms[j] = acc
    comb = np.empty(nl + 1, f8)
    comb[:nl] = Fnd + Fsm[:nl]
    comb[nl] = F_intern - F_net[0]
    for j in range(nl):
        if marked_red[j] == 1:
            comb[nl] = F_intern - F_net[j + 1]
            break
    T_old = T.copy()
    for i in range(nl + 1):
        c = comb[i]
        if itervalue == 0:
            pref[i] = 1e-2
        if itervalue == 6000:
            pref[i] = 1e-3
        delta_t = pref[i] * play[0] / abs(c) ** 0.5 if c != 0 else 0.0
        dT = c / (pint[0] - pint[1]) * delta_t
        if abs(dT) > 20.0:
            dT = 20.0 * c / abs(c)
        if itervalue % adapt_interval == 0:
            T_store[i] = T_old[i]
        if itervalue % adapt_interval == adapt_interval - 1:
            if abs(T_old[i] - T_store[i]) < adapt_interval / 2.0 * abs(dT):
                pref[i] /= 1.5
            else:
                pref[i] *= 1.1
        T[i] = max(T_old[i] + dT, 1.001)
    return dict(tlay=T, T_store=T_store, prefactor=pref, F_net_diff=Fnd, F_smooth=Fsm, F_smooth_sum=Fsms)


# ------------------------------------------------------------------------------- post-processing
def integrate_optdepth_transmission_iso(trans_wg, delta_tau_wg, gauss_weight, nbin, nlayer, ny):
    """K:2888-2912 -> (trans_band, delta_tau_band)"""
    gw = 0.5 * np.asarray(gauss_weight, f8)
    tr = np.asarray(trans_wg, f8)[:nlayer * nbin * ny].reshape(nlayer, nbin, ny)
    dt = np.asarray(delta_tau_wg, f8)[:nlayer * nbin * ny].reshape(nlayer, nbin, ny)
    tb = np.zeros((nlayer, nbin))
    db = np.zeros((nlayer, nbin))
    for y in range(ny):
        db += gw[y] * dt[:, :, y]
        tb += gw[y] * tr[:, :, y]
    return tb.reshape(-1), db.reshape(-1)


def integrate_optdepth_transmission_noniso(tr_u, tr_l, dt_u, dt_l, gauss_weight, dtc_u, dtc_l, nbin, nlayer, ny):
    """K:2916-2947 -> (trans_band, delta_tau_band, delta_tau_all_clouds)"""
    gw = 0.5 * np.asarray(gauss_weight, f8)
    r = lambda a: np.asarray(a, f8)[:nlayer * nbin * ny].reshape(nlayer, nbin, ny)
    tu, tl, du, dl = r(tr_u), r(tr_l), r(dt_u), r(dt_l)
    tb = np.zeros((nlayer, nbin))
    db = np.zeros((nlayer, nbin))
    for y in range(ny):
        db += gw[y] * (du[:, :, y] + dl[:, :, y])
        tb += gw[y] * (tu[:, :, y] * tl[:, :, y])
    dtc = np.asarray(dtc_l, f8)[:nlayer * nbin] + np.asarray(dtc_u, f8)[:nlayer * nbin]
    return tb.reshape(-1), db.reshape(-1), dtc


def calc_contr_func(trans_a, trans_b, trans_weight_band, gauss_weight, planckband_lay, epsi, nbin, nlayer, ny):
    """K:2951-2983 (trans_b None) / K:2987-3020 -> (trans_weight_band, contr_func_band)"""
    gw = 0.5 * np.asarray(gauss_weight, f8)
    t = np.asarray(trans_a, f8)[:nlayer * nbin * ny].reshape(nlayer, nbin, ny)
    if trans_b is not None:
        t = t * np.asarray(trans_b, f8)[:nlayer * nbin * ny].reshape(nlayer, nbin, ny)
    tw = np.array(trans_weight_band, f8)[:nlayer * nbin].reshape(nlayer, nbin)
    for i in range(nlayer):
        for y in range(ny):
            to_top = np.ones(nbin)
            for j in range(i + 1, nlayer):
                to_top = to_top * t[j, :, y]
            tw[i] += gw[y] * (1.0 - t[i, :, y]) * to_top
    B = np.asarray(planckband_lay, f8).reshape(nbin, nlayer + 2)
    cf = 2.0 * PI * epsi * B[:, :nlayer].T * tw
    return tw.reshape(-1), cf.reshape(-1)


def dB_dT(lam, T):
    """K:294-308"""
    D = 2.0 * HCONST * CSPEED ** 3 * HCONST / (lam ** 6 * KBOLTZMANN * (T * T))
    with np.errstate(over="ignore", invalid="ignore"):
        ex = np.exp(HCONST * CSPEED / (lam * KBOLTZMANN * T))
        return D * ex / ((ex - 1.0) * (ex - 1.0))


def integrated_dB_dT(kw, ky, lbot, ltop, T):
    """K:312-329"""
    r = np.zeros_like(lbot)
    for y in range(len(ky)):
        x = (ky[y] - 0.5) * 2.0
        arg = (ltop - lbot) / 2.0 * x + (ltop + lbot) / 2.0
        r = r + (ltop - lbot) / 2.0 * kw[y] * dB_dT(arg, T)
    return r


def calc_mean_opacities(opac_wg_lay, abs_cl_lay, meanmolmass_lay, planckband_lay, interwave, deltawave, T_lay,
                        gauss_weight, gauss_y, nlayer, nbin, ny, T_star):
    """K:3024-3115 -> dict planck_T_pl, ross_T_pl, planck_T_star, ross_T_star ([i]), opac_band ([i][x])"""
    gw = np.asarray(gauss_weight, f8)
    gy = np.asarray(gauss_y, f8)
    k = np.asarray(opac_wg_lay, f8)[:nlayer * nbin * ny].reshape(nlayer, nbin, ny)
    band = np.zeros((nlayer, nbin))
    for y in range(ny):
        band += 0.5 * gw[y] * k[:, :, y]
    B = np.asarray(planckband_lay, f8).reshape(nbin, nlayer + 2)
    iw = np.asarray(interwave, f8)
    dw = np.asarray(deltawave, f8)
    out = {q: np.zeros(nlayer) for q in ("planck_T_pl", "ross_T_pl", "planck_T_star", "ross_T_star")}
    with np.errstate(invalid="ignore", divide="ignore"):
        for i in range(nlayer):
            kk = band[i] + np.asarray(abs_cl_lay, f8)[i * nbin:(i + 1) * nbin] / meanmolmass_lay[i]
            dpl = integrated_dB_dT(gw, gy, iw[:-1], iw[1:], T_lay[i])
            dst = integrated_dB_dT(gw, gy, iw[:-1], iw[1:], T_star)
            pos = kk > 0
            out["planck_T_pl"][i] = np.sum(kk * B[:, i] * dw) / np.sum(B[:, i] * dw)
            out["ross_T_pl"][i] = np.sum(dpl) / np.sum(np.where(pos, dpl / kk, 0.0))
            out["planck_T_star"][i] = np.sum(kk * B[:, nlayer] * dw) / np.sum(B[:, nlayer] * dw)
            out["ross_T_star"][i] = np.sum(dst) / np.sum(np.where(pos, dst / kk, 0.0))
            if T_lay[i] < 70:
                out["ross_T_pl"][i] = -3
            if T_star < 70:
                out["planck_T_star"][i] = -3
                out["ross_T_star"][i] = -3
    out["opac_band"] = band.reshape(-1)
    return out


def integrate_beamflux(F_dir_band, deltalambda, nbin, nint):
    """K:3119-3139"""
    Fb = np.asarray(F_dir_band, f8)[:nint * nbin].reshape(nint, nbin)
    tot = np.zeros(nint)
    for x in range(nbin):
        tot += Fb[:, x] * deltalambda[x]
    return tot
