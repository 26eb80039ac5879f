// Two-stream flux sweeps: iterative (fband_iso / fband_noniso) and Thomas-matrix variants.
// From-scratch sm_100a kernels for K:1366-2424 of the reference.
//
// Mapping.  The reference gives thread (x, y) the column y + ny*x, so neighbouring threads are ny
// doubles apart in every [i][x][y] array.  Here one thread owns the FLAT column c = y + ny*x: a warp
// reads/writes 32 consecutive doubles of every array at every layer (fully coalesced), and the layer
// recursion lives in registers.  Columns are independent, therefore the reference's (3*scat+1) or
// (1000*scat+1) back-to-back launches (C:531-537) collapse into an in-kernel loop over passes: pass
// p+1 of a column reads only what the same thread wrote in pass p.
//
// Scheduling.  The grid is persistent (a multiple of the SM count); each block walks over tiles of
// FB_THREADS columns and finishes ALL passes of a tile before moving on, so that from the second pass
// on the coefficient slab of the tile (about 7 kB per column) is served by the 126 MB L2 and HBM sees
// the coefficients once per flux solve instead of twice per pass.
#include "common.cuh"
#include "sweep_math.cuh"
#include "fband_plan.cuh"

#define FB_THREADS 128

struct FbandScalars {
    double g_0, Rstar, a, f_factor, mu_star, epsi, delta_tau_limit, i2s_transition;
    int nint, nbin, ny, dir_beam, clouds, scat_corr, npass;
};

__device__ __forceinline__ double tiny_abs(double f) { return fabs(f) < 1e-100 ? fabs(f) : f; }

// ------------------------------------------------------------------------------------------------
// isothermal layers (K:1366-1517)
// ------------------------------------------------------------------------------------------------
template <bool CLOUDS, bool SCORR>
__global__ void __launch_bounds__(FB_THREADS)
k_fband_iso(double* __restrict__ F_down, double* __restrict__ F_up, const double* __restrict__ F_dir,
            const double* __restrict__ planck, const double* __restrict__ w_0, const double* __restrict__ Mt,
            const double* __restrict__ Nt, const double* __restrict__ Pt, const double* __restrict__ Gp,
            const double* __restrict__ Gm, const double* __restrict__ albedo,
            const double* __restrict__ g0tot, FbandScalars s) {
    const int nint = s.nint, nlay = nint - 1;
    const int ncol = s.nbin * s.ny;
    const int ntile = (ncol + FB_THREADS - 1) / FB_THREADS;
    const double neg_mu = -s.mu_star;

    for (int tile = blockIdx.x; tile < ntile; tile += gridDim.x) {
        const int col = tile * FB_THREADS + threadIdx.x;
        if (col >= ncol) continue;
        const int x = col / s.ny;
        const double* __restrict__ B = planck + (size_t)x * (nlay + 2);  // [x][i], i fastest
        const double A_s = albedo[x];
        const double toa = (1.0 - s.dir_beam) * s.f_factor * ((s.Rstar / s.a) * (s.Rstar / s.a)) * hc::PI * B[nlay];
        const double B_surf = B[nlay + 1];

        for (int pass = 0; pass < s.npass; pass++) {
            double w0 = 0.0, E = 1.0;
            // ---- downward sweep, TOA -> BOA ----
            double Fd = toa;
            F_down[col + (size_t)ncol * nlay] = Fd;
            double Fdir_above = F_dir[col + (size_t)ncol * nlay];
#pragma unroll 4
            for (int i = nlay - 1; i >= 0; i--) {
                const size_t e = col + (size_t)ncol * i;
                w0 = w_0[e];
                const double M = Mt[e], N = Nt[e], P = Pt[e], G_pl = Gp[e], G_min = Gm[e];
                const double Fdir_i = F_dir[e];
                const double Fup_i = F_up[e];
                const double g0 = CLOUDS ? g0tot[x + (size_t)s.nbin * i] : s.g_0;
                E = SCORR ? E_parameter(w0, g0, s.i2s_transition) : 1.0;
                const double direct_terms = beam_source(Fdir_i, Fdir_above, neg_mu, M, G_min, N, G_pl, P, G_min);
                Fd = tiny_to_abs(sweep_update(1.0 / M, P, N, Fd, Fup_i, source_factor(s.epsi, w0, E),
                                              planck_iso(B[i], M, N, P), direct_terms));
                F_down[e] = Fd;
                Fdir_above = Fdir_i;
            }
            // ---- upward sweep, BOA -> TOA.  w0/E still hold layer 0 (K:1472) ----
            double Fdir_below = Fdir_above;  // F_dir at interface 0
            double Fu = boa_flux(A_s, Fdir_below, Fd, w0, E, B_surf);
            F_up[col] = Fu;
#pragma unroll 4
            for (int i = 1; i < nint; i++) {
                const size_t el = col + (size_t)ncol * (i - 1);  // layer below interface i
                const size_t e = el + ncol;
                w0 = w_0[el];
                const double M = Mt[el], N = Nt[el], P = Pt[el], G_pl = Gp[el], G_min = Gm[el];
                const double Fdir_i = F_dir[e];
                const double Fd_i = F_down[e];
                const double g0 = CLOUDS ? g0tot[x + (size_t)s.nbin * (i - 1)] : s.g_0;
                E = SCORR ? E_parameter(w0, g0, s.i2s_transition) : 1.0;
                const double direct_terms = beam_source(Fdir_i, Fdir_below, neg_mu, N, G_min, M, G_pl, P, G_pl);
                Fu = tiny_to_abs(sweep_update(1.0 / M, P, N, Fu, Fd_i, source_factor(s.epsi, w0, E),
                                              planck_iso(B[i - 1], M, N, P), direct_terms));
                F_up[e] = Fu;
                Fdir_below = Fdir_i;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// non-isothermal layers (K:1521-1799): every layer is split into a lower and an upper half with a
// linear-in-tau Planck source; falls back to the isothermal form when the half-layer is optically thin.
// ------------------------------------------------------------------------------------------------
struct NonisoCoef {
    const double *w0_u, *w0_l, *dtau_u, *dtau_l, *dtc_u, *dtc_l, *M_u, *M_l, *N_u, *N_l, *P_u, *P_l, *Gp_u,
        *Gp_l, *Gm_u, *Gm_l;
};

struct HalfLayer {
    double w0, dtau, M, N, P, Gp, Gm, g0, E;
};

template <bool CLOUDS, bool SCORR>
__device__ __forceinline__ void load_halves(const NonisoCoef& c, const double* __restrict__ g0_lay,
                                            const double* __restrict__ g0_int, size_t e, size_t b, int nbin,
                                            const FbandScalars& s, HalfLayer& up, HalfLayer& low) {
    up.w0 = c.w0_u[e];   low.w0 = c.w0_l[e];
    up.dtau = c.dtau_u[e] + c.dtc_u[b];
    low.dtau = c.dtau_l[e] + c.dtc_l[b];
    up.M = c.M_u[e];     low.M = c.M_l[e];
    up.N = c.N_u[e];     low.N = c.N_l[e];
    up.P = c.P_u[e];     low.P = c.P_l[e];
    up.Gp = c.Gp_u[e];   low.Gp = c.Gp_l[e];
    up.Gm = c.Gm_u[e];   low.Gm = c.Gm_l[e];
    up.g0 = s.g_0;       low.g0 = s.g_0;
    if (CLOUDS) {
        const double gl = g0_lay[b];
        up.g0 = (gl + g0_int[b + nbin]) / 2.0;
        low.g0 = (g0_int[b] + gl) / 2.0;
    }
    up.E = SCORR ? E_parameter(up.w0, up.g0, s.i2s_transition) : 1.0;
    low.E = SCORR ? E_parameter(low.w0, low.g0, s.i2s_transition) : 1.0;
}

template <bool CLOUDS, bool SCORR>
__global__ void __launch_bounds__(FB_THREADS)
k_fband_noniso(double* __restrict__ F_down, double* __restrict__ F_up, double* __restrict__ Fc_down,
               double* __restrict__ Fc_up, const double* __restrict__ F_dir, const double* __restrict__ Fc_dir,
               const double* __restrict__ planck_lay, const double* __restrict__ planck_int, NonisoCoef c,
               const double* __restrict__ albedo, const double* __restrict__ g0_lay,
               const double* __restrict__ g0_int, FbandScalars s) {
    const int nint = s.nint, nlay = nint - 1;
    const int ncol = s.nbin * s.ny;
    const int ntile = (ncol + FB_THREADS - 1) / FB_THREADS;
    const double neg_mu = -s.mu_star;

    for (int tile = blockIdx.x; tile < ntile; tile += gridDim.x) {
        const int col = tile * FB_THREADS + threadIdx.x;
        if (col >= ncol) continue;
        const int x = col / s.ny;
        const double* __restrict__ BL = planck_lay + (size_t)x * (nlay + 2);
        const double* __restrict__ BI = planck_int + (size_t)x * nint;
        const double A_s = albedo[x];
        const double toa = (1.0 - s.dir_beam) * s.f_factor * ((s.Rstar / s.a) * (s.Rstar / s.a)) * hc::PI * BL[nlay];
        const double B_surf = BL[nlay + 1];

        for (int pass = 0; pass < s.npass; pass++) {
            HalfLayer up, low;
            low.w0 = 0.0;
            low.E = 1.0;
            // ---- downward sweep ----
            double Fd = toa;
            F_down[col + (size_t)ncol * nlay] = Fd;
            double Fdir_above = F_dir[col + (size_t)ncol * nlay];
            double Bint_above = BI[nlay];
#pragma unroll 2
            for (int i = nlay - 1; i >= 0; i--) {
                const size_t e = col + (size_t)ncol * i;
                const size_t b = (size_t)x + (size_t)s.nbin * i;
                load_halves<CLOUDS, SCORR>(c, g0_lay, g0_int, e, b, s.nbin, s, up, low);
                const double Blay = BL[i], Bint = BI[i];
                const double Fdir_i = F_dir[e], Fcdir_i = Fc_dir[e];
                // upper half: interface i+1 -> layer centre (K:1640-1664)
                double pt = up.dtau < s.delta_tau_limit
                                ? planck_thin(Bint_above, Blay, up.M, up.N, up.P)
                                : planck_grad_down(Blay, Bint_above, up.M, up.N, up.P,
                                                   gradient_factor(s.epsi, up.w0, up.g0, up.E),
                                                   __ddiv_rn(__dsub_rn(Blay, Bint_above), up.dtau));
                double dr = beam_source(Fcdir_i, Fdir_above, neg_mu, up.M, up.Gm, up.N, up.Gp, up.Gm, up.P);
                const double Fc = tiny_to_abs(sweep_update(1.0 / up.M, up.P, up.N, Fd, Fc_up[e],
                                                           source_factor(s.epsi, up.w0, up.E), pt, dr));
                Fc_down[e] = Fc;
                // lower half: layer centre -> interface i (K:1667-1691)
                pt = low.dtau < s.delta_tau_limit
                         ? planck_thin(Bint, Blay, low.M, low.N, low.P)
                         : planck_grad_down(Bint, Blay, low.M, low.N, low.P,
                                            gradient_factor(s.epsi, low.w0, low.g0, low.E),
                                            __ddiv_rn(__dsub_rn(Bint, Blay), low.dtau));
                dr = beam_source(Fdir_i, Fcdir_i, neg_mu, low.M, low.Gm, low.N, low.Gp, low.P, low.Gm);
                Fd = tiny_to_abs(sweep_update(1.0 / low.M, low.P, low.N, Fc, F_up[e],
                                              source_factor(s.epsi, low.w0, low.E), pt, dr));
                F_down[e] = Fd;
                Fdir_above = Fdir_i;
                Bint_above = Bint;
            }
            // ---- upward sweep; low.w0 / low.E still hold the lower half of layer 0 (K:1704) ----
            double Fdir_below = Fdir_above;
            double Fu = boa_flux(A_s, Fdir_below, Fd, low.w0, low.E, B_surf);
            F_up[col] = Fu;
            double Bint_below = BI[0];
#pragma unroll 2
            for (int i = 1; i < nint; i++) {
                const size_t el = col + (size_t)ncol * (i - 1);
                const size_t e = el + ncol;
                const size_t b = (size_t)x + (size_t)s.nbin * (i - 1);
                load_halves<CLOUDS, SCORR>(c, g0_lay, g0_int, el, b, s.nbin, s, up, low);
                const double Blay = BL[i - 1], Bint = BI[i];
                const double Fdir_i = F_dir[e], Fcdir_l = Fc_dir[el];
                // lower half: interface i-1 -> layer centre (K:1744-1768)
                double pt = low.dtau < s.delta_tau_limit
                                ? planck_thin(Bint_below, Blay, low.M, low.N, low.P)
                                : planck_grad_up(Blay, Bint_below, low.M, low.N, low.P,
                                                 gradient_factor(s.epsi, low.w0, low.g0, low.E),
                                                 __ddiv_rn(__dsub_rn(Bint_below, Blay), low.dtau));
                double dr = beam_source(Fcdir_l, Fdir_below, neg_mu, low.N, low.Gm, low.M, low.Gp, low.P, low.Gp);
                // no tiny-value clean-up here: the reference applies it to index i instead of i-1 (K:1763)
                const double Fcu = sweep_update(1.0 / low.M, low.P, low.N, Fu, Fc_down[el],
                                                source_factor(s.epsi, low.w0, low.E), pt, dr);
                Fc_up[el] = Fcu;
                // upper half: layer centre -> interface i (K:1771-1795)
                pt = up.dtau < s.delta_tau_limit
                         ? planck_thin(Bint, Blay, up.M, up.N, up.P)
                         : planck_grad_up(Bint, Blay, up.M, up.N, up.P,
                                          gradient_factor(s.epsi, up.w0, up.g0, up.E),
                                          __ddiv_rn(__dsub_rn(Blay, Bint), up.dtau));
                dr = beam_source(Fdir_i, Fcdir_l, neg_mu, up.N, up.Gm, up.M, up.Gp, up.P, up.Gp);
                Fu = tiny_to_abs(sweep_update(1.0 / up.M, up.P, up.N, Fcu, F_down[e],
                                              source_factor(s.epsi, up.w0, up.E), pt, dr));
                F_up[e] = Fu;
                Fdir_below = Fdir_i;
                Bint_below = Bint;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Thomas-matrix variants (K:1803-2424).  Unknown vector per column:
//   iso:    [Fd_0, Fu_0, Fd_1, Fu_1, ...]                      (2*nint unknowns)
//   noniso: [Fd_0, Fu_0, Fcd_0, Fcu_0, Fd_1, Fu_1, ...]        (4*nint-2 unknowns)
// The reference first stores alpha/beta/source terms for every (half-)layer in four global scratch
// arrays and then eliminates; here they are formed on the fly inside the elimination loop (each
// (half-)layer feeds two consecutive matrix rows), so only c'/d' touch memory.
// ------------------------------------------------------------------------------------------------
struct RowTerms {
    double alpha, beta, src_down, src_up;
};

template <bool CLOUDS, bool SCORR>
__global__ void __launch_bounds__(FB_THREADS)
k_fband_matrix_iso(double* __restrict__ F_down, double* __restrict__ F_up, const double* __restrict__ F_dir,
                   const double* __restrict__ planck, const double* __restrict__ w_0,
                   const double* __restrict__ Mt, const double* __restrict__ Nt, const double* __restrict__ Pt,
                   const double* __restrict__ Gp, const double* __restrict__ Gm,
                   const double* __restrict__ g0tot, double* __restrict__ c_prime, double* __restrict__ d_prime,
                   const int* __restrict__ scat_trigger, const double* __restrict__ trans_wg,
                   const double* __restrict__ albedo, FbandScalars s) {
    const int nint = s.nint, nlay = nint - 1;
    const int ncol = s.nbin * s.ny;
    const int col = blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= ncol) return;
    const int x = col / s.ny;
    const double* __restrict__ B = planck + (size_t)x * (nlay + 2);
    const double A_s = albedo[x];
    const double neg_mu = -s.mu_star;
    const double two_pi_eps = 2.0 * hc::PI * s.epsi;
    const double toa = (1.0 - s.dir_beam) * s.f_factor * ((s.Rstar / s.a) * (s.Rstar / s.a)) * hc::PI * B[nlay];

    if (scat_trigger[col] == 1) {
        const int n_matrix = 2 * nint;
        // bottom boundary row (K:1902-1924)
        double w0 = w_0[col];
        double g0 = CLOUDS ? g0tot[x] : s.g_0;
        double E = SCORR ? E_parameter(w0, g0, s.i2s_transition) : 1.0;
        const double src_boa = A_s * F_dir[col] + (1.0 - A_s) * hc::PI * (1.0 - w0) / (E - w0) * B[nlay + 1];
        double b_i = -A_s, c_i = 1.0, d_i = src_boa;
        double cp = c_i / b_i, dp = d_i / b_i;
        c_prime[col] = cp;
        d_prime[col] = dp;
        RowTerms r{0, 0, 0, 0};
        for (int i = 1; i < n_matrix - 1; i++) {
            const double c_prev = c_i;
            if (i & 1) {
                const int j = (i - 1) >> 1;  // new layer: form its terms (K:1866-1895)
                const size_t e = col + (size_t)ncol * j;
                w0 = w_0[e];
                const double M = Mt[e], N = Nt[e], P = Pt[e], G_pl = Gp[e], G_min = Gm[e];
                if (CLOUDS) g0 = g0tot[x + (size_t)s.nbin * j];
                if (SCORR) E = E_parameter(w0, g0, s.i2s_transition);
                r.alpha = P / M;
                r.beta = -N / M;
                const double planck_terms = two_pi_eps * (1.0 - w0) / (E - w0) * (N + M - P) * B[j];
                const double Fdj = F_dir[e], Fdj1 = F_dir[e + ncol];
                double dd = Fdj / neg_mu * (G_min * M + G_pl * N) - Fdj1 / neg_mu * P * G_min;
                dd = fmin(0.0, dd);
                r.src_down = 1.0 / M * (planck_terms + dd);
                double du = Fdj1 / neg_mu * (G_min * N + G_pl * M) - Fdj / neg_mu * P * G_pl;
                du = fmin(0.0, du);
                r.src_up = 1.0 / M * (planck_terms + du);
                b_i = -r.beta; c_i = -r.alpha; d_i = r.src_down;
            } else {
                b_i = -r.beta; c_i = 1.0; d_i = r.src_up;
            }
            const double den = b_i - c_prev * cp;
            const double ncp = c_i / den;
            const double ndp = (d_i - c_prev * dp) / den;
            cp = ncp; dp = ndp;
            c_prime[col + (size_t)ncol * i] = cp;
            d_prime[col + (size_t)ncol * i] = dp;
        }
        // top boundary row (K:1946-1950)
        double x_i = (toa - c_i * dp) / (0.0 - c_i * cp);
        d_prime[col + (size_t)ncol * (n_matrix - 1)] = x_i;
        F_up[col + (size_t)ncol * nlay] = x_i;
        for (int i = n_matrix - 2; i >= 0; i--) {
            x_i = d_prime[col + (size_t)ncol * i] - c_prime[col + (size_t)ncol * i] * x_i;
            if ((i & 1) == 0) F_down[col + (size_t)ncol * (i >> 1)] = x_i;
            else F_up[col + (size_t)ncol * ((i - 1) >> 1)] = x_i;
        }
    } else {
        // pure absorption (K:1969-2022)
        double Fd = toa;
        F_down[col + (size_t)ncol * nlay] = Fd;
        for (int i = nlay - 1; i >= 0; i--) {
            const size_t e = col + (size_t)ncol * i;
            const double trans = trans_wg[e];
            Fd = trans * Fd + two_pi_eps * (1.0 - trans) * B[i];
            Fd = tiny_abs(Fd);
            F_down[e] = Fd;
        }
        double Fu = A_s * (F_dir[col] + Fd) + (1.0 - A_s) * hc::PI * B[nlay + 1];
        F_up[col] = Fu;
        for (int i = 1; i < nint; i++) {
            const size_t el = col + (size_t)ncol * (i - 1);
            const double trans = trans_wg[el];
            Fu = trans * Fu + two_pi_eps * (1.0 - trans) * B[i - 1];
            Fu = tiny_abs(Fu);
            F_up[el + ncol] = Fu;
        }
    }
}

template <bool CLOUDS, bool SCORR>
__global__ void __launch_bounds__(FB_THREADS)
k_fband_matrix_noniso(double* __restrict__ F_down, double* __restrict__ F_up, double* __restrict__ Fc_down,
                      double* __restrict__ Fc_up, const double* __restrict__ F_dir,
                      const double* __restrict__ Fc_dir, const double* __restrict__ planck_lay,
                      const double* __restrict__ planck_int, NonisoCoef c, const double* __restrict__ g0_lay,
                      const double* __restrict__ g0_int, double* __restrict__ c_prime,
                      double* __restrict__ d_prime, const int* __restrict__ scat_trigger,
                      const double* __restrict__ trans_u, const double* __restrict__ trans_l,
                      const double* __restrict__ albedo, FbandScalars s) {
    const int nint = s.nint, nlay = nint - 1;
    const int ncol = s.nbin * s.ny;
    const int col = blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= ncol) return;
    const int x = col / s.ny;
    const double* __restrict__ BL = planck_lay + (size_t)x * (nlay + 2);
    const double* __restrict__ BI = planck_int + (size_t)x * nint;
    const double A_s = albedo[x];
    const double neg_mu = -s.mu_star;
    const double two_pi_eps = 2.0 * hc::PI * s.epsi;
    const double toa = (1.0 - s.dir_beam) * s.f_factor * ((s.Rstar / s.a) * (s.Rstar / s.a)) * hc::PI * BL[nlay];

    if (scat_trigger[col] == 1) {
        const int n_matrix = 4 * nint - 2;
        double w0 = c.w0_l[col];
        double g0 = s.g_0;
        if (CLOUDS) g0 = (g0_int[x] + g0_lay[x]) / 2.0;
        double E = SCORR ? E_parameter(w0, g0, s.i2s_transition) : 1.0;
        const double src_boa = A_s * F_dir[col] + (1.0 - A_s) * hc::PI * (1.0 - w0) / (E - w0) * BL[nlay + 1];
        double b_i = -A_s, c_i = 1.0, d_i = src_boa;
        double cp = c_i / b_i, dp = d_i / b_i;
        c_prime[col] = cp;
        d_prime[col] = dp;
        RowTerms r{0, 0, 0, 0};
        for (int i = 1; i < n_matrix - 1; i++) {
            const double c_prev = c_i;
            if (i & 1) {
                const int j = (i - 1) >> 1;  // half-layer index: even = lower half, odd = upper half
                const int L = j >> 1;
                const size_t e = col + (size_t)ncol * L;
                const size_t b = (size_t)x + (size_t)s.nbin * L;
                double M, N, P, G_min, G_pl, del_tau, pt_down, pt_up, dd, du;
                if ((j & 1) == 0) {  // lower half (K:2111-2149)
                    M = c.M_l[e]; N = c.N_l[e]; P = c.P_l[e]; w0 = c.w0_l[e];
                    G_min = c.Gm_l[e]; G_pl = c.Gp_l[e];
                    del_tau = c.dtau_l[e] + c.dtc_l[b];
                    if (CLOUDS) g0 = (g0_int[b] + g0_lay[b]) / 2.0;
                    if (SCORR) E = E_parameter(w0, g0, s.i2s_transition);
                    const double Bi = BI[L], Bl = BL[L];
                    if (del_tau < s.delta_tau_limit) {
                        pt_up = (N + M - P) * (Bi + Bl) / 2.0;
                        pt_down = pt_up;
                    } else {
                        const double pgrad = (Bi - Bl) / del_tau;
                        pt_down = (M + N) * Bi - P * Bl + s.epsi / (E * (1.0 - w0 * g0)) * (P - M + N) * pgrad;
                        pt_up = (M + N) * Bl - P * Bi + s.epsi / (E * (1.0 - w0 * g0)) * (M - N - P) * pgrad;
                    }
                    const double Fd0 = F_dir[e], Fc0 = Fc_dir[e];
                    dd = Fd0 / neg_mu * (G_min * M + G_pl * N) - Fc0 / neg_mu * P * G_min;
                    du = Fc0 / neg_mu * (G_min * N + G_pl * M) - Fd0 / neg_mu * P * G_pl;
                } else {  // upper half (K:2150-2188)
                    M = c.M_u[e]; N = c.N_u[e]; P = c.P_u[e]; w0 = c.w0_u[e];
                    G_min = c.Gm_u[e]; G_pl = c.Gp_u[e];
                    del_tau = c.dtau_u[e] + c.dtc_u[b];
                    if (CLOUDS) g0 = (g0_int[b + s.nbin] + g0_lay[b]) / 2.0;
                    if (SCORR) E = E_parameter(w0, g0, s.i2s_transition);
                    const double Bi = BI[L + 1], Bl = BL[L];
                    if (del_tau < s.delta_tau_limit) {
                        pt_up = (N + M - P) * (Bl + Bi) / 2.0;
                        pt_down = pt_up;
                    } else {
                        const double pgrad = (Bl - Bi) / del_tau;
                        pt_down = (M + N) * Bl - P * Bi + s.epsi / (E * (1.0 - w0 * g0)) * (P - M + N) * pgrad;
                        pt_up = (M + N) * Bi - P * Bl + s.epsi / (E * (1.0 - w0 * g0)) * (M - N - P) * pgrad;
                    }
                    const double Fd1 = F_dir[e + ncol], Fc0 = Fc_dir[e];
                    dd = Fc0 / neg_mu * (G_min * M + G_pl * N) - Fd1 / neg_mu * P * G_min;
                    du = Fd1 / neg_mu * (G_min * N + G_pl * M) - Fc0 / neg_mu * P * G_pl;
                }
                dd = fmin(0.0, dd);
                du = fmin(0.0, du);
                r.alpha = P / M;
                r.beta = -N / M;
                const double pre = two_pi_eps * (1.0 - w0) / (E - w0);
                r.src_down = 1.0 / M * (pre * pt_down + dd);
                r.src_up = 1.0 / M * (pre * pt_up + du);
                b_i = -r.beta; c_i = -r.alpha; d_i = r.src_down;
            } else {
                b_i = -r.beta; c_i = 1.0; d_i = r.src_up;
            }
            const double den = b_i - c_prev * cp;
            const double ncp = c_i / den;
            const double ndp = (d_i - c_prev * dp) / den;
            cp = ncp; dp = ndp;
            c_prime[col + (size_t)ncol * i] = cp;
            d_prime[col + (size_t)ncol * i] = dp;
        }
        double x_i = (toa - c_i * dp) / (0.0 - c_i * cp);
        d_prime[col + (size_t)ncol * (n_matrix - 1)] = x_i;
        F_up[col + (size_t)ncol * nlay] = x_i;
        for (int i = n_matrix - 2; i >= 0; i--) {
            x_i = d_prime[col + (size_t)ncol * i] - c_prime[col + (size_t)ncol * i] * x_i;
            if (x_i < 1e-100) x_i = fabs(x_i);  // K:2264 (flips every negative value)
            const size_t o = col + (size_t)ncol * (i >> 2);
            switch (i & 3) {
                case 0: F_down[o] = x_i; break;
                case 1: F_up[o] = x_i; break;
                case 2: Fc_down[o] = x_i; break;
                default: Fc_up[o] = x_i; break;
            }
        }
    } else {
        // pure absorption with half layers (K:2286-2422)
        double Fd = toa;
        F_down[col + (size_t)ncol * nlay] = Fd;
        for (int i = nlay - 1; i >= 0; i--) {
            const size_t e = col + (size_t)ncol * i;
            const size_t b = (size_t)x + (size_t)s.nbin * i;
            const double tu = trans_u[e], tl = trans_l[e];
            const double dtu = c.dtau_u[e] + c.dtc_u[b], dtl = c.dtau_l[e] + c.dtc_l[b];
            const double Bl = BL[i], Bi1 = BI[i + 1], Bi = BI[i];
            double pt;
            if (dtu < s.delta_tau_limit) pt = (Bi1 + Bl) / 2.0 * (1.0 - tu);
            else pt = Bl - tu * Bi1 + s.epsi * (tu - 1.0) * ((Bl - Bi1) / dtu);
            double Fc = tu * Fd + two_pi_eps * pt;
            Fc = tiny_abs(Fc);
            Fc_down[e] = Fc;
            if (dtl < s.delta_tau_limit) pt = (Bi + Bl) / 2.0 * (1.0 - tl);
            else pt = Bi - tl * Bl + s.epsi * (tl - 1.0) * ((Bi - Bl) / dtl);
            Fd = tl * Fc + two_pi_eps * pt;
            Fd = tiny_abs(Fd);
            F_down[e] = Fd;
        }
        double Fu = A_s * (F_dir[col] + Fd) + (1.0 - A_s) * hc::PI * BL[nlay + 1];
        F_up[col] = Fu;
        for (int i = 1; i < nint; i++) {
            const size_t el = col + (size_t)ncol * (i - 1);
            const size_t b = (size_t)x + (size_t)s.nbin * (i - 1);
            const double tu = trans_u[el], tl = trans_l[el];
            const double dtu = c.dtau_u[el] + c.dtc_u[b], dtl = c.dtau_l[el] + c.dtc_l[b];
            const double Bl = BL[i - 1], Bi0 = BI[i - 1], Bi = BI[i];
            double pt;
            if (dtl < s.delta_tau_limit) pt = (Bi0 + Bl) / 2.0 * (1.0 - tl);
            else pt = Bl - tl * Bi0 + s.epsi * ((Bi0 - Bl) / dtl) * (1.0 - tl);
            const double Fcu = tl * Fu + two_pi_eps * pt;  // no tiny_abs: K:2394 indexes i, not i-1
            Fc_up[el] = Fcu;
            if (dtu < s.delta_tau_limit) pt = (Bi + Bl) / 2.0 * (1.0 - tu);
            else pt = Bi - tu * Bl + s.epsi * ((Bl - Bi) / dtu) * (1.0 - tu);
            Fu = tu * Fcu + two_pi_eps * pt;
            Fu = tiny_abs(Fu);
            F_up[el + ncol] = Fu;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// layer-parallel variants (fband_cp.cu)
struct CpNonisoCoef {
    const double *w0_u, *w0_l, *dtau_u, *dtau_l, *dtc_u, *dtc_l, *M_u, *M_l, *N_u, *N_l, *P_u, *P_l, *Gp_u,
        *Gp_l, *Gm_u, *Gm_l;
};
int fband_iso_cp_try(helios_ctx* ctx, double* F_down, double* F_up, const double* F_dir, const double* planck,
                     const double* w_0, const double* M, const double* N, const double* P, const double* Gp,
                     const double* Gm, const double* albedo, const double* g0tot, double g_0, double Rstar,
                     double a, int nint, int nbin, double f_factor, double mu_star, int ny, double epsi,
                     int dir_beam, int clouds, int scat_corr, double i2s, int npass);
int fband_noniso_cp_try(helios_ctx* ctx, double* F_down, double* F_up, double* Fc_down, double* Fc_up,
                        const double* F_dir, const double* Fc_dir, const double* planck_lay,
                        const double* planck_int, CpNonisoCoef c, const double* albedo, const double* g0_lay,
                        const double* g0_int, double g_0, double Rstar, double a, int nint, int nbin,
                        double f_factor, double mu_star, int ny, double epsi, double delta_tau_limit,
                        int dir_beam, int clouds, int scat_corr, double i2s, int npass);


static int fband_grid(helios_ctx* ctx, int ncol) {
    const int ntile = (ncol + FB_THREADS - 1) / FB_THREADS;
    // persistent grid: at most 8 resident 128-thread blocks per SM
    const int cap = ctx->num_sms * 8;
    return ntile < cap ? ntile : cap;
}

#define DISPATCH2(KERNEL, clouds, scorr, ...)                                   \
    do {                                                                        \
        if (clouds) {                                                           \
            if (scorr) KERNEL<true, true> __VA_ARGS__;                          \
            else KERNEL<true, false> __VA_ARGS__;                               \
        } else {                                                                \
            if (scorr) KERNEL<false, true> __VA_ARGS__;                         \
            else KERNEL<false, false> __VA_ARGS__;                              \
        }                                                                       \
    } while (0)

extern "C" {

int helios_fband_iso(helios_ctx* ctx, double* F_down_wg, double* F_up_wg, const double* F_dir_wg,
                     const double* planckband_lay, const double* w_0, const double* M_term,
                     const double* N_term, const double* P_term, const double* G_plus,
                     const double* G_minus, const double* surf_albedo, const double* g_0_tot_lay,
                     double g_0, int singlewalk, double Rstar, double a, int numinterfaces, int nbin,
                     double f_factor, double mu_star, int ny, double epsi, int dir_beam, int clouds,
                     int scat_corr, int debug, double i2s_transition, int npass) {
    HCTX(ctx);
    (void)singlewalk; (void)debug;
    HARG(F_down_wg && F_up_wg && F_dir_wg && planckband_lay && w_0 && M_term && N_term && P_term &&
         G_plus && G_minus && surf_albedo);
    HARG(clouds == 0 || g_0_tot_lay != nullptr);
    HARG(numinterfaces > 1 && nbin > 0 && ny > 0 && npass > 0);
    HBATCHDIMS(ctx, numinterfaces == ctx->batch.nint() && nbin == ctx->batch.nbin && ny == ctx->batch.ny);
    if (ctx->fband_mode != 1) {
        const int rc = fband_iso_cp_try(ctx, F_down_wg, F_up_wg, F_dir_wg, planckband_lay, w_0, M_term, N_term,
                                        P_term, G_plus, G_minus, surf_albedo, g_0_tot_lay, g_0, Rstar, a,
                                        numinterfaces, nbin, f_factor, mu_star, ny, epsi, dir_beam, clouds == 1,
                                        scat_corr == 1, i2s_transition, npass);
        if (rc >= 0) return rc;
        if (ctx->fband_mode == 2) {
            helios_set_error("helios_fband_iso: shape does not fit the layer-parallel kernel");
            return HELIOS_ERR_ARG;
        }
    }
    HNOBATCH(ctx);  // only the layer-parallel kernel has a batched form
    FbandScalars s{g_0, Rstar, a, f_factor, mu_star, epsi, 0.0, i2s_transition,
                   numinterfaces, nbin, ny, dir_beam, clouds, scat_corr, npass};
    const int grid = fband_grid(ctx, nbin * ny);
    DISPATCH2(k_fband_iso, clouds == 1, scat_corr == 1,
              <<<grid, FB_THREADS, 0, ctx->stream>>>(F_down_wg, F_up_wg, F_dir_wg, planckband_lay, w_0,
                                                     M_term, N_term, P_term, G_plus, G_minus, surf_albedo,
                                                     g_0_tot_lay, s));
    HLAUNCHED(ctx);
    return HELIOS_OK;
}

int helios_fband_noniso(
    helios_ctx* ctx, double* F_down_wg, double* F_up_wg, double* Fc_down_wg, double* Fc_up_wg,
    const double* F_dir_wg, const double* Fc_dir_wg, const double* planckband_lay,
    const double* planckband_int, const double* w_0_upper, const double* w_0_lower,
    const double* delta_tau_wg_upper, const double* delta_tau_wg_lower,
    const double* delta_tau_all_clouds_upper, const double* delta_tau_all_clouds_lower,
    const double* M_upper, const double* M_lower, const double* N_upper, const double* N_lower,
    const double* P_upper, const double* P_lower, const double* G_plus_upper, const double* G_plus_lower,
    const double* G_minus_upper, const double* G_minus_lower, const double* surf_albedo,
    const double* g_0_tot_lay, const double* g_0_tot_int, double g_0, int singlewalk, double Rstar,
    double a, int numinterfaces, int nbin, double f_factor, double mu_star, int ny, double epsi,
    double delta_tau_limit, int dir_beam, int clouds, int scat_corr, int debug, double i2s_transition,
    int npass) {
    HCTX(ctx);
    (void)singlewalk; (void)debug;
    HARG(F_down_wg && F_up_wg && Fc_down_wg && Fc_up_wg && F_dir_wg && Fc_dir_wg && planckband_lay &&
         planckband_int && w_0_upper && w_0_lower && delta_tau_wg_upper && delta_tau_wg_lower &&
         delta_tau_all_clouds_upper && delta_tau_all_clouds_lower && M_upper && M_lower && N_upper &&
         N_lower && P_upper && P_lower && G_plus_upper && G_plus_lower && G_minus_upper && G_minus_lower &&
         surf_albedo);
    HARG(clouds == 0 || (g_0_tot_lay != nullptr && g_0_tot_int != nullptr));
    HARG(numinterfaces > 1 && nbin > 0 && ny > 0 && npass > 0);
    HBATCHDIMS(ctx, numinterfaces == ctx->batch.nint() && nbin == ctx->batch.nbin && ny == ctx->batch.ny);
    FbandScalars s{g_0, Rstar, a, f_factor, mu_star, epsi, delta_tau_limit, i2s_transition,
                   numinterfaces, nbin, ny, dir_beam, clouds, scat_corr, npass};
    NonisoCoef c{w_0_upper, w_0_lower, delta_tau_wg_upper, delta_tau_wg_lower, delta_tau_all_clouds_upper,
                 delta_tau_all_clouds_lower, M_upper, M_lower, N_upper, N_lower, P_upper, P_lower,
                 G_plus_upper, G_plus_lower, G_minus_upper, G_minus_lower};
    if (ctx->fband_mode != 1) {
        CpNonisoCoef cc{w_0_upper, w_0_lower, delta_tau_wg_upper, delta_tau_wg_lower, delta_tau_all_clouds_upper,
                        delta_tau_all_clouds_lower, M_upper, M_lower, N_upper, N_lower, P_upper, P_lower,
                        G_plus_upper, G_plus_lower, G_minus_upper, G_minus_lower};
        const int rc = fband_noniso_cp_try(ctx, F_down_wg, F_up_wg, Fc_down_wg, Fc_up_wg, F_dir_wg, Fc_dir_wg,
                                           planckband_lay, planckband_int, cc, surf_albedo, g_0_tot_lay,
                                           g_0_tot_int, g_0, Rstar, a, numinterfaces, nbin, f_factor, mu_star, ny,
                                           epsi, delta_tau_limit, dir_beam, clouds == 1, scat_corr == 1,
                                           i2s_transition, npass);
        if (rc >= 0) return rc;
        if (ctx->fband_mode == 2) {
            helios_set_error("helios_fband_noniso: shape does not fit the layer-parallel kernel");
            return HELIOS_ERR_ARG;
        }
    }
    HNOBATCH(ctx);  // only the layer-parallel kernel has a batched form
    const int grid = fband_grid(ctx, nbin * ny);
    DISPATCH2(k_fband_noniso, clouds == 1, scat_corr == 1,
              <<<grid, FB_THREADS, 0, ctx->stream>>>(F_down_wg, F_up_wg, Fc_down_wg, Fc_up_wg, F_dir_wg,
                                                     Fc_dir_wg, planckband_lay, planckband_int, c,
                                                     surf_albedo, g_0_tot_lay, g_0_tot_int, s));
    HLAUNCHED(ctx);
    return HELIOS_OK;
}

int helios_fband_noniso_plan_build(
    helios_ctx* ctx, double* plan, const double* F_dir_wg, const double* Fc_dir_wg, const double* w_0_upper,
    const double* w_0_lower, const double* delta_tau_wg_upper, const double* delta_tau_wg_lower,
    const double* delta_tau_all_clouds_upper, const double* delta_tau_all_clouds_lower, const double* M_upper,
    const double* M_lower, const double* N_upper, const double* N_lower, const double* P_upper, const double* P_lower,
    const double* G_plus_upper, const double* G_plus_lower, const double* G_minus_upper, const double* G_minus_lower,
    const double* surf_albedo, const double* g_0_tot_lay, const double* g_0_tot_int, double g_0, int numinterfaces,
    int nbin, double mu_star, int ny, double epsi, double delta_tau_limit, int clouds, int scat_corr,
    double i2s_transition) {
    HCTX(ctx);
    HARG(plan && F_dir_wg && Fc_dir_wg && w_0_upper && w_0_lower && delta_tau_wg_upper && delta_tau_wg_lower &&
         delta_tau_all_clouds_upper && delta_tau_all_clouds_lower && M_upper && M_lower && N_upper && N_lower &&
         P_upper && P_lower && G_plus_upper && G_plus_lower && G_minus_upper && G_minus_lower && surf_albedo);
    HARG(clouds == 0 || (g_0_tot_lay != nullptr && g_0_tot_int != nullptr));
    HARG(numinterfaces > 1 && nbin > 0 && ny > 0);
    HBATCHDIMS(ctx, numinterfaces == ctx->batch.nint() && nbin == ctx->batch.nbin && ny == ctx->batch.ny);
    const double* coef[16] = {w_0_upper, w_0_lower, delta_tau_wg_upper, delta_tau_wg_lower, delta_tau_all_clouds_upper,
                              delta_tau_all_clouds_lower, M_upper, M_lower, N_upper, N_lower, P_upper, P_lower,
                              G_plus_upper, G_plus_lower, G_minus_upper, G_minus_lower};
    // dir_beam is not an argument of this entry point: the beam rows are dropped only when both beam arrays are
    // known to hold zeros (fdir_noniso ran with dir_beam == 0 and nothing wrote to them since)
    const int rc = plan2_noniso_build(ctx, plan, F_dir_wg, Fc_dir_wg, coef, surf_albedo, g_0_tot_lay, g_0_tot_int, g_0,
                                      mu_star, epsi, delta_tau_limit, numinterfaces, nbin, ny, 0, clouds == 1,
                                      scat_corr == 1, i2s_transition);
    if (rc < 0) {
        helios_set_error("helios_fband_noniso_plan_build: more than 128 layers are not supported by the planned sweep");
        return HELIOS_ERR_ARG;
    }
    return rc;
}

int helios_fband_noniso_plan_size(helios_ctx* ctx, int numinterfaces, int nbin, int ny, size_t* ndoubles) {
    HCTX(ctx);
    HARG(ndoubles != nullptr && numinterfaces > 1 && nbin > 0 && ny > 0);
    *ndoubles = plan2_noniso_size(numinterfaces, nbin * ny, ctx->batch.nbatch);
    return HELIOS_OK;
}

int helios_fband_noniso_planned(helios_ctx* ctx, double* F_down_wg, double* F_up_wg, double* Fc_down_wg,
                                double* Fc_up_wg, const double* plan, const double* planckband_lay,
                                const double* planckband_int, const double* surf_albedo, double Rstar, double a,
                                int numinterfaces, int nbin, double f_factor, int ny, int dir_beam, int npass) {
    HCTX(ctx);
    HARG(F_down_wg && F_up_wg && Fc_down_wg && Fc_up_wg && plan && planckband_lay && planckband_int && surf_albedo);
    HARG(numinterfaces > 1 && nbin > 0 && ny > 0 && npass > 0);
    HBATCHDIMS(ctx, numinterfaces == ctx->batch.nint() && nbin == ctx->batch.nbin && ny == ctx->batch.ny);
    const int rc = plan2_noniso_sweep(ctx, F_down_wg, F_up_wg, Fc_down_wg, Fc_up_wg, plan, planckband_lay, planckband_int,
                                      surf_albedo, Rstar, a, numinterfaces, nbin, f_factor, ny, dir_beam, npass);
    if (rc == -2) {
        helios_set_error("helios_fband_noniso_planned: `plan` was not built by helios_fband_noniso_plan_build for this shape "
                         "(or was overwritten since)");
        return HELIOS_ERR_STATE;
    }
    if (rc < 0) {
        helios_set_error("helios_fband_noniso_planned: more than 128 layers are not supported by the planned sweep");
        return HELIOS_ERR_ARG;
    }
    return rc;
}

int helios_fband_iso_plan_size(helios_ctx* ctx, int numinterfaces, int nbin, int ny, size_t* ndoubles) {
    HCTX(ctx);
    HARG(ndoubles != nullptr && numinterfaces > 1 && nbin > 0 && ny > 0);
    *ndoubles = plan2_iso_size(numinterfaces, nbin * ny, ctx->batch.nbatch);
    return HELIOS_OK;
}

int helios_fband_iso_plan_build(helios_ctx* ctx, double* plan, const double* F_dir_wg, const double* w_0,
                                const double* M_term, const double* N_term, const double* P_term,
                                const double* G_plus, const double* G_minus, const double* surf_albedo,
                                const double* g_0_tot_lay, double g_0, int numinterfaces, int nbin, double mu_star,
                                int ny, double epsi, int dir_beam, int clouds, int scat_corr, double i2s_transition) {
    HCTX(ctx);
    HARG(plan && F_dir_wg && w_0 && M_term && N_term && P_term && G_plus && G_minus && surf_albedo);
    HARG(clouds == 0 || g_0_tot_lay != nullptr);
    HARG(numinterfaces > 1 && nbin > 0 && ny > 0);
    HBATCHDIMS(ctx, numinterfaces == ctx->batch.nint() && nbin == ctx->batch.nbin && ny == ctx->batch.ny);
    const int rc = plan2_iso_build(ctx, plan, F_dir_wg, w_0, M_term, N_term, P_term, G_plus, G_minus, surf_albedo,
                                   g_0_tot_lay, g_0, mu_star, epsi, numinterfaces, nbin, ny, dir_beam, clouds == 1,
                                   scat_corr == 1, i2s_transition);
    if (rc < 0) {
        helios_set_error("helios_fband_iso_plan_build: more than 256 layers are not supported by the planned sweep");
        return HELIOS_ERR_ARG;
    }
    return rc;
}

int helios_fband_iso_planned(helios_ctx* ctx, double* F_down_wg, double* F_up_wg, const double* plan,
                             const double* planckband_lay, const double* surf_albedo, double Rstar, double a,
                             int numinterfaces, int nbin, double f_factor, int ny, int dir_beam, int npass) {
    HCTX(ctx);
    HARG(F_down_wg && F_up_wg && plan && planckband_lay && surf_albedo);
    HARG(numinterfaces > 1 && nbin > 0 && ny > 0 && npass > 0);
    HBATCHDIMS(ctx, numinterfaces == ctx->batch.nint() && nbin == ctx->batch.nbin && ny == ctx->batch.ny);
    const int rc = plan2_iso_sweep(ctx, F_down_wg, F_up_wg, plan, planckband_lay, surf_albedo, Rstar, a, numinterfaces,
                                   nbin, f_factor, ny, dir_beam, npass);
    if (rc == -2) {
        helios_set_error("helios_fband_iso_planned: `plan` was not built by helios_fband_iso_plan_build for this shape (or "
                         "was overwritten since)");
        return HELIOS_ERR_STATE;
    }
    if (rc < 0) {
        helios_set_error("helios_fband_iso_planned: more than 256 layers are not supported by the planned sweep");
        return HELIOS_ERR_ARG;
    }
    return rc;
}

int helios_fband_matrix_iso(
    helios_ctx* ctx, double* F_down_wg, double* F_up_wg, const double* F_dir_wg,
    const double* planckband_lay, const double* w_0, const double* M_term, const double* N_term,
    const double* P_term, const double* G_plus, const double* G_minus, const double* g_0_tot_lay,
    double* alpha, double* beta, double* source_term_down, double* source_term_up, double* c_prime,
    double* d_prime, const int* scat_trigger, const double* trans_wg, const double* surf_albedo,
    double g_0, int singlewalk, double Rstar, double a, int numinterfaces, int nbin, double f_factor,
    double mu_star, int ny, double epsi, int dir_beam, int clouds, int scat_corr, int debug,
    double i2s_transition) {
    HCTX(ctx);
    HNOBATCH(ctx);
    (void)singlewalk; (void)debug; (void)alpha; (void)beta; (void)source_term_down; (void)source_term_up;
    HARG(F_down_wg && F_up_wg && F_dir_wg && planckband_lay && w_0 && M_term && N_term && P_term &&
         G_plus && G_minus && c_prime && d_prime && scat_trigger && trans_wg && surf_albedo);
    HARG(clouds == 0 || g_0_tot_lay != nullptr);
    HARG(numinterfaces > 1 && nbin > 0 && ny > 0);
    FbandScalars s{g_0, Rstar, a, f_factor, mu_star, epsi, 0.0, i2s_transition,
                   numinterfaces, nbin, ny, dir_beam, clouds, scat_corr, 1};
    const int grid = ceil_div((long long)nbin * ny, FB_THREADS);
    DISPATCH2(k_fband_matrix_iso, clouds == 1, scat_corr == 1,
              <<<grid, FB_THREADS, 0, ctx->stream>>>(F_down_wg, F_up_wg, F_dir_wg, planckband_lay, w_0,
                                                     M_term, N_term, P_term, G_plus, G_minus, g_0_tot_lay,
                                                     c_prime, d_prime, scat_trigger, trans_wg,
                                                     surf_albedo, s));
    HLAUNCHED(ctx);
    return HELIOS_OK;
}

int helios_fband_matrix_noniso(
    helios_ctx* ctx, double* F_down_wg, double* F_up_wg, double* Fc_down_wg, double* Fc_up_wg,
    const double* F_dir_wg, const double* Fc_dir_wg, const double* planckband_lay,
    const double* planckband_int, const double* w_0_upper, const double* w_0_lower,
    const double* delta_tau_wg_upper, const double* delta_tau_wg_lower,
    const double* delta_tau_all_clouds_upper, const double* delta_tau_all_clouds_lower,
    const double* M_upper, const double* M_lower, const double* N_upper, const double* N_lower,
    const double* P_upper, const double* P_lower, const double* G_plus_upper, const double* G_plus_lower,
    const double* G_minus_upper, const double* G_minus_lower, const double* g_0_tot_lay,
    const double* g_0_tot_int, double* alpha, double* beta, double* source_term_down,
    double* source_term_up, double* c_prime, double* d_prime, const int* scat_trigger,
    const double* trans_wg_upper, const double* trans_wg_lower, const double* surf_albedo, double g_0,
    int singlewalk, double Rstar, double a, int numinterfaces, int nbin, double f_factor, double mu_star,
    int ny, double epsi, double delta_tau_limit, int dir_beam, int clouds, int scat_corr, int debug,
    double i2s_transition) {
    HCTX(ctx);
    HNOBATCH(ctx);
    (void)singlewalk; (void)debug; (void)alpha; (void)beta; (void)source_term_down; (void)source_term_up;
    HARG(F_down_wg && F_up_wg && Fc_down_wg && Fc_up_wg && F_dir_wg && Fc_dir_wg && planckband_lay &&
         planckband_int && w_0_upper && w_0_lower && delta_tau_wg_upper && delta_tau_wg_lower &&
         delta_tau_all_clouds_upper && delta_tau_all_clouds_lower && M_upper && M_lower && N_upper &&
         N_lower && P_upper && P_lower && G_plus_upper && G_plus_lower && G_minus_upper && G_minus_lower &&
         c_prime && d_prime && scat_trigger && trans_wg_upper && trans_wg_lower && surf_albedo);
    HARG(clouds == 0 || (g_0_tot_lay != nullptr && g_0_tot_int != nullptr));
    HARG(numinterfaces > 1 && nbin > 0 && ny > 0);
    FbandScalars s{g_0, Rstar, a, f_factor, mu_star, epsi, delta_tau_limit, i2s_transition,
                   numinterfaces, nbin, ny, dir_beam, clouds, scat_corr, 1};
    NonisoCoef c{w_0_upper, w_0_lower, delta_tau_wg_upper, delta_tau_wg_lower, delta_tau_all_clouds_upper,
                 delta_tau_all_clouds_lower, M_upper, M_lower, N_upper, N_lower, P_upper, P_lower,
                 G_plus_upper, G_plus_lower, G_minus_upper, G_minus_lower};
    const int grid = ceil_div((long long)nbin * ny, FB_THREADS);
    DISPATCH2(k_fband_matrix_noniso, clouds == 1, scat_corr == 1,
              <<<grid, FB_THREADS, 0, ctx->stream>>>(F_down_wg, F_up_wg, Fc_down_wg, Fc_up_wg, F_dir_wg,
                                                     Fc_dir_wg, planckband_lay, planckband_int, c,
                                                     g_0_tot_lay, g_0_tot_int, c_prime, d_prime,
                                                     scat_trigger, trans_wg_upper, trans_wg_lower,
                                                     surf_albedo, s));
    HLAUNCHED(ctx);
    return HELIOS_OK;
}

}  // extern "C"
